#!/bin/bash
python -m pytest tests/test_gpu_parity.py -q -x -k "synth30" 2>&1 | tail -3
ANM_SOLVER=sparse python -m pytest tests/test_gpu_parity.py -q -x -k "golden or batch_vs_oracle or radial_tree or synth30" 2>&1 | tail -3
python tools/bench_synth30.py 8192 | tail -1
ANM_SOLVER=generic python tools/bench_synth30.py 8192 | tail -1
