#!/bin/bash
# One GPU-box pass: tests, smoke, bench (both arms), ncu launch list + full capture of the rollout kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 300 python tools/bench_synth30.py > gpurun_out/synth30.log 2>&1; tail -2 gpurun_out/synth30.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:anm_env_kernel -c 1 -f \
    -o gpurun_out/prof python tools/rollout_ncu.py 4096 50 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
