#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "mixed or batch_vs_oracle or golden_trajectory" 2>&1 | tail -5
for f in 0 1e-2 1e-3 1e-4; do
  python bench.py --no-cpu-baseline --steps 6000 --solver-fp32-above $f 2>>gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fp32_above=$f value %.4g  ms/step %.4f  e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
done
