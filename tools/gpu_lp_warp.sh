#!/bin/bash
# Round 2, warp-per-program LP kernel: its GPU tests, config 5 with it, launch list (in that order: budget is short).
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_zz_lp_gpu.py -x -q -k warp > gpurun_out/r02_lp_warp_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_lp_warp_tests.log
tail -4 gpurun_out/r02_lp_warp_tests.log
ANM_LP_KERNEL=warp timeout 45 python tools/bench_config5.py --lp gpu --steps 64 > gpurun_out/r02_config5_gpu_lp_warp_1gpu.json 2> gpurun_out/r02_config5_gpu_lp_warp_1gpu.err; echo "bench rc=$?"
tail -c 900 gpurun_out/r02_config5_gpu_lp_warp_1gpu.json; tail -3 gpurun_out/r02_config5_gpu_lp_warp_1gpu.err
ANM_LP_KERNEL=warp timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_config5_warp_launches.csv python tools/bench_config5.py --lp gpu --steps 4 > gpurun_out/r02_config5_warp_ncu.log 2>&1; echo "ncu rc=$?"
grep lp_solve gpurun_out/r02_config5_warp_launches.csv | awk -F'","' '{print $NF}' | head -8
