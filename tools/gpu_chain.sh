#!/bin/bash
# Launch-chaining check: the new parity tests, then bench with and without programmatic dependent launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout or chained or queued or chaining or host_buffer or autoreset or golden_trajectory" 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; echo "bench rc=$?"; cat gpurun_out/bench_chain.json; tail -5 gpurun_out/bench_chain.err
ANM_PDL=0 timeout 600 python bench.py --no-cpu-baseline --steps 4000 > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench(nopdl) rc=$?"; cat gpurun_out/bench_nopdl.json; tail -5 gpurun_out/bench_nopdl.err
