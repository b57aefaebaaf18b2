#!/usr/bin/env python
"""Single-warp latency of the Newton loop: a small batch of instances that ALL diverge
(100 iterations each, reference semantics), timed with CUDA events; profile it with
    ncu --set full --import-source on -k regex:anm_env_kernel -s 5 -c 1 -o gpurun_out/strag python tools/straggler_bench.py
and read per-line samples with tools/ncu_by_line.py (every sample is straggler time)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.env_spec import HostEnvSpec  # noqa: E402
from gym_anm_b200.native import NativeBatch  # noqa: E402
from gym_anm_b200.networks import anm6_network  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spec = HostEnvSpec(anm6_network(), "state", 0, 0.25, 0.9, 100)
nb = NativeBatch(spec, B)
# inputs of a divergent golden transition (tests/golden/transitions_anm6.npz, step 7)
p_load = np.tile([-8.39346436, -21.6534211, -23.70692933], (B, 1))
p_pot = np.tile([32.21884094, 51.46035968], (B, 1))
p_set = np.tile([9.26470233, 21.16009022, -22.28164456], (B, 1))
q_set = np.tile([17.74864205, -55.19842769, -51.90732803], (B, 1))
args = [torch.as_tensor(a, device="cuda") for a in (p_load, p_pot, p_set, q_set)]
soc = torch.full((B, 1), 0.5, dtype=torch.float64, device="cuda")
for _ in range(3):
    nb.set_state(soc=soc)
    full, r, e, pe, conv = nb.transition(*args)
torch.cuda.synchronize()
assert not bool(conv.any()), "expected every instance to diverge"
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
ev0.record()
for _ in range(N):
    nb.transition(*args)
ev1.record()
torch.cuda.synchronize()
us = ev0.elapsed_time(ev1) * 1000 / N
print("B=%d all-divergent transition: %.1f us per launch -> %.2f us per Newton iteration (100 iterations)" % (B, us, us / 100))
