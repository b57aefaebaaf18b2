#!/bin/bash
# Round 2, batched LP solver: the whole GPU suite (the LP tests sort last), config 5 with the LPs on the GPU, launch list.
mkdir -p gpurun_out
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/r02_lp_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_lp_tests.log
tail -5 gpurun_out/r02_lp_tests.log
timeout 120 python tools/bench_config5.py --lp gpu --steps 64 > gpurun_out/r02_config5_gpu_lp_1gpu.json 2> gpurun_out/r02_config5_gpu_lp_1gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_config5_gpu_lp_1gpu.json; tail -3 gpurun_out/r02_config5_gpu_lp_1gpu.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_config5_launches.csv python tools/bench_config5.py --lp gpu --steps 6 > gpurun_out/r02_config5_ncu.log 2>&1; echo "ncu rc=$?"
grep -c lp_solve_kernel gpurun_out/r02_config5_launches.csv
