#!/bin/bash
mkdir -p gpurun_out
for lib in gym_anm_b200/lib/libanm_b200_t*.so; do
  ANM_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 6000 2>>gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
done
ANM_B200_LIB=$PWD/gym_anm_b200/lib/libanm_b200_t64m7.so python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
