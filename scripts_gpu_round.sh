#!/bin/bash
# One GPU-box pass: smoke, bench (both arms), ncu launch list + full capture of the step kernel.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:anm_env_kernel -s 30 -c 3 -f -o gpurun_out/prof \
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
