#!/bin/bash
# A/B of the Newton back-ends on the ANM6Easy bench + parity tests with the radial solver forced
for sv in dense radial; do
  ANM_SOLVER=$sv python bench.py --no-cpu-baseline --steps 6000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ANM_SOLVER=$sv value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
done
ANM_SOLVER=radial python -m pytest tests/test_gpu_parity.py -q -x -k "golden or batch_vs_oracle or radial_tree or full_size or seeded" 2>&1 | tail -3
