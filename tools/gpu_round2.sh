#!/bin/bash
# round-2 quick pass on one GPU: tests, the bench line (driver arguments), an A/B of variant libraries
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
show='import sys,json; d=json.loads(sys.stdin.read()); print("%s: value %.4g us/step %.2f | isolated %.4g | chained %.4g lockstep %.4g | e2e %.4g sync %.4g | hbm frac %.4f fp64 frac %s" % (sys.argv[1], d["value"], 1000*d["ms_per_step"], d["isolated_block"]["value"], d["chained"]["value"], d["lockstep"]["value"], d["e2e"]["value"], d["sync_every_step"]["value"], d["roofline"]["frac"], d["roofline"]["fp64"] and round(d["roofline"]["fp64"]["frac"],4)))'
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_k20.err
python -c "$show" k20 < gpurun_out/r2_bench_k20.json
for lib in gym_anm_b200/lib/libanm_b200_*.so; do
  [ -f "$lib" ] || continue
  ANM_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" $lib
done
python bench.py --no-cpu-baseline 2>/dev/null | python -c "$show" k2000
