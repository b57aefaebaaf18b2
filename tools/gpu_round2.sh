#!/bin/bash
# Round-2 full pass on one GPU: tests, smoke, the bench lines (driver arguments, reference arm, config 4, config-3 size),
# the ncu launch list of the bench command and one `--set full` capture of a 1000-step rollout launch (outputs > L2).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 600 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_config4.json 2> gpurun_out/r02_bench_config4.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/r02_bench_config4.json; tail -2 gpurun_out/r02_bench_config4.err
timeout 600 python bench.py --envs 8192 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_8192.json 2> /dev/null; echo "8192 rc=$?"; cut -c1-200 gpurun_out/r02_bench_8192.json
timeout 600 python bench.py --envs 16384 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_16384.json 2> /dev/null; echo "16384 rc=$?"; cut -c1-200 gpurun_out/r02_bench_16384.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:anm_env_kernel -c 1 -f \
    -o gpurun_out/r02_prof python tools/rollout_ncu.py 4096 1000 > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu2 rc=$?"; tail -2 gpurun_out/r02_ncu_full.log
