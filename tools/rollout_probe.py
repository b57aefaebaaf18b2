#!/usr/bin/env python
"""us per step of anm_rollout (one launch, T steps) for a few (B, T); device-resident actions."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

for B, T in ((4096, 200), (256, 200), (16384, 100)):
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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for r in range(4):
        nb.rollout(acts, out=out, chained=(r > 0))
    ev1.record()
    torch.cuda.synchronize()
    us = 1000 * ev0.elapsed_time(ev1) / (4 * T)
    print("rollout B=%5d T=%4d: %.2f us/step, %.3g env-steps/s, terminated %.2f %%" % (B, T, us, B / us * 1e6, 100 * float(out[2].double().mean())))
    del env, nb, acts, out
