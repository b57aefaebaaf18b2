#!/usr/bin/env python
"""One profiled anm_step launch (the closed-loop unit: every instance advances by ONE step, fully ordered launch).
Run under `ncu --profile-from-start off`: shows how a lock-step launch spends its time once the ~1 % of instances
whose Newton iteration runs to the cap are the only ones left (profiles/r02_lockstep_launch.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedANM6Easy(B, validate_actions=False)
nb = env.native
env.reset(seed=3)
nb.set_autoreset_pool(env.state.clone())
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
lo, hi = (torch.as_tensor(x, device="cuda") for x in (env.spec.action_low, env.spec.action_high))
acts = torch.rand((24, B, 6), dtype=torch.float64, device="cuda", generator=gen) * (hi - lo) + lo
out = (nb.empty(B, 18), nb.empty(B), nb.empty(B, dtype=torch.uint8))
nit = torch.zeros(B, dtype=torch.int32, device="cuda")
for t in range(20):
    nb.step(acts[t], None, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for t in range(20, 23):
    nb.step(acts[t], None, out=out, extras={"n_iter": nit})
    torch.cuda.synchronize()
    print("step", t, "instances at the iteration cap:", int((nit >= 100).sum()), flush=True)
torch.cuda.cudart().cudaProfilerStop()
