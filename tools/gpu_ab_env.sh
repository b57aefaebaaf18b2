#!/bin/bash
# generic A/B over environment switches: tests once, then the bench line per variant.
# usage: gpu_ab_env.sh [--notest] name1:VAR=val,VAR2=val name2: ...
mkdir -p gpurun_out
if [ "$1" != "--notest" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6; else shift; fi
show='import sys,json; d=json.loads(sys.stdin.read()); print("%s: value %.4g us/step %.2f | isolated %.4g | chained %.4g lockstep %.4g | e2e %.4g sync %.4g | fp64 frac %s" % (sys.argv[1], d["value"], 1000*d["ms_per_step"], d["isolated_block"]["value"], d["chained"]["value"], d["lockstep"]["value"], d["e2e"]["value"], d["sync_every_step"]["value"], d["roofline"]["fp64"] and round(d["roofline"]["fp64"]["frac"],4)))'
for spec in "$@"; do
  name=${spec%%:*}; vars=${spec#*:}
  envs=$(echo "$vars" | tr ',' ' ')
  env $envs timeout 400 python bench.py --steps ${STEPS:-2000} --no-cpu-baseline 2>gpurun_out/ab_$name.err | python -c "$show" $name || tail -3 gpurun_out/ab_$name.err
done
