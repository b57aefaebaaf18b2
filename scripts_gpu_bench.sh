#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do python bench.py --no-cpu-baseline --steps 10000 2>gpurun_out/bench_quick.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"; done
