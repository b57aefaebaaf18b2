#!/bin/bash
# A/B: round-1 tree (_r1tmp, old bench) vs the current library and its no-guard variant; old bench format
show='import sys,json; d=json.loads(sys.stdin.read()); print("%s: value %.4g  us/step %.2f  chained %.4g  lockstep %.4g  e2e %.4g" % (sys.argv[1], d["value"], 1000*d["ms_per_step"], d["per_step_launches"]["chained"]["value"], d["per_step_launches"]["lockstep"]["value"], d["e2e"]["value"]))'
(cd _r1tmp && python bench.py --no-cpu-baseline --steps 4000 2>/dev/null | python -c "$show" r1)
cp _r1tmp/bench.py /tmp/old_bench.py
for lib in gym_anm_b200/lib/libanm_b200.so gym_anm_b200/lib/libanm_b200_noguard.so; do
  ANM_B200_LIB=$PWD/$lib python _r1tmp/bench_cur.py --no-cpu-baseline --steps 4000 2>/dev/null | python -c "$show" $lib
done
