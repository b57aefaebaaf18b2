#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:anm_env_kernel -s 30 -c 2 -f -o gpurun_out/prof2 \
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches2.csv \
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1; echo "ncu1 rc=$?"
