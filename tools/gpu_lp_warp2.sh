#!/bin/bash
# Round 2, warp-per-program LP kernel (8-row evaluation groups, batched pivot update): tests, config 5, one ncu capture.
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_zz_lp_gpu.py -x -q -k warp > gpurun_out/r02_lp_warp2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_lp_warp2_tests.log
tail -3 gpurun_out/r02_lp_warp2_tests.log
timeout 40 python tools/bench_config5.py --lp gpu --steps 64 > gpurun_out/r02_config5_gpu_lp_warp2_1gpu.json 2> gpurun_out/r02_config5_gpu_lp_warp2_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_config5_gpu_lp_warp2_1gpu.json"))
print({k: d[k] for k in ("lp_kernel", "value", "lp_s_per_step", "env_step_s_per_step", "first_solve_s", "mean_pivots_per_solve", "lp_stats", "oracle_check")})
PY
timeout 45 ncu --set full --clock-control none --import-source on -k regex:lp_solve_warp --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_lp_warp python tools/bench_config5.py --lp gpu --steps 3 > gpurun_out/r02_lp_warp_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_lp_warp.ncu-rep 2>/dev/null
