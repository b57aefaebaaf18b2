#!/usr/bin/env python
"""One profiled anm_rollout launch (run under `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
env = BatchedANM6Easy(B, validate_actions=False)
nb = env.native
env.reset(seed=3)
nb.set_autoreset_pool(env.state.clone())
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
lo, hi = (torch.as_tensor(x, device="cuda") for x in (env.spec.action_low, env.spec.action_high))
acts = torch.rand((T, B, 6), dtype=torch.float64, device="cuda", generator=gen) * (hi - lo) + lo
out = (nb.empty(T, B, 18), nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8))
nb.rollout(acts, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
nb.rollout(acts, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
