#!/bin/bash
# A/B of the host-buffer I/O modes of anm_step_host (e2e leg of bench.py)
for m in copy zc_out zc; do
  ANM_HOST_IO=$m python bench.py --no-cpu-baseline --steps 6000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ANM_HOST_IO=$m value %.4g  e2e %.4g (%.1f us/step)'%(d['value'], d['e2e']['value'], 4096/d['e2e']['value']*1e6))"
done
python -m pytest tests/test_gpu_parity.py -q -x -k host_buffer 2>&1 | tail -2
ANM_HOST_IO=zc python -m pytest tests/test_gpu_parity.py -q -x -k host_buffer 2>&1 | tail -2
