#!/bin/bash
# A/B: the in-tree library vs variant builds gym_anm_b200/lib/libanm_b200_*.so (bench value + parity subset)
python bench.py --no-cpu-baseline --steps 6000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default lib: value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
for lib in gym_anm_b200/lib/libanm_b200_*.so; do
  ANM_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 6000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib: value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
  ANM_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -q -x -k "golden or batch_vs_oracle or radial_tree or seeded" 2>&1 | tail -1
done
