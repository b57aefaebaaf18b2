#!/bin/bash
# A/B: the in-tree library vs variant builds gym_anm_b200/lib/libanm_b200_*.so (bench value + parity subset)
show='import sys,json; d=json.loads(sys.stdin.read()); print("%s: value %.4g  us/step %.2f  chained %.4g  lockstep %.4g  e2e %.4g" % (sys.argv[1], d["value"], 1000*d["ms_per_step"], d["per_step_launches"]["chained"]["value"], d["per_step_launches"]["lockstep"]["value"], d["e2e"]["value"]))'
python bench.py --no-cpu-baseline --steps 8000 2>/dev/null | python -c "$show" default
for lib in gym_anm_b200/lib/libanm_b200_*.so; do
  ANM_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 8000 2>/dev/null | python -c "$show" $lib
  ANM_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -q -x -k "golden_trajectory or batch_vs_oracle or rollout_equals" 2>&1 | tail -1
done
