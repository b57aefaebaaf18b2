#!/usr/bin/env python
"""us per step of anm_rollout (one launch, T steps) for a few (B, T); device-resident actions.

`scale` shrinks the uniform-random actions around the middle of the action box: 1.0 = the bench's random agent (most
set-points infeasible, ~0.9 % divergent solves), 0.0 = a "do nothing" policy (every set-point interior: the projection
leaves after its first trip, no divergent solves) -- the two ends between which a trained policy lives."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

for B, T, scale in ((4096, 200, 1.0), (4096, 200, 0.3), (4096, 200, 0.0), (256, 200, 1.0), (16384, 100, 1.0)):
    env = BatchedANM6Easy(B, validate_actions=False)
    nb = env.native
    env.reset(seed=3)
    nb.set_autoreset_pool(env.state.clone())
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    lo, hi = (torch.as_tensor(x, device="cuda") for x in (env.spec.action_low, env.spec.action_high))
    mid = torch.as_tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0], device="cuda", dtype=torch.float64)
    acts = (torch.rand((T, B, 6), dtype=torch.float64, device="cuda", generator=gen) * (hi - lo) + lo - mid) * scale + mid
    out = (nb.empty(T, B, 18), nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8))
    nb.rollout(acts, out=out)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for r in range(4):
        nb.rollout(acts, out=out, chained=(r > 0))
    ev1.record()
    torch.cuda.synchronize()
    us = 1000 * ev0.elapsed_time(ev1) / (4 * T)
    print("rollout B=%5d T=%4d action scale %.1f: %.2f us/step, %.3g env-steps/s, terminated %.2f %%" % (B, T, scale, us, B / us * 1e6, 100 * float(out[2].double().mean())))
    del env, nb, acts, out
