#!/bin/bash
# A/B of kernel variants on ONE box: value (graph replay) only, short runs.
mkdir -p gpurun_out
for lib in gym_anm_b200/lib/libanm_b200_*.so; do
  ANM_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 4000 2>>gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
done
