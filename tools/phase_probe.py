#!/usr/bin/env python
"""SM cycles per phase of one step launch (diagnostic build: ANM_B200_LIB=.../libanm_b200_diag.so)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedANM6Easy(B, validate_actions=False)
env.reset(seed=2020)
nb = env.native
nb.set_autoreset_pool(env.state.clone())
rng = np.random.default_rng(0)
stats = torch.zeros(B * 20, dtype=torch.int32, device="cuda")
obs, rew, term = nb.empty(B, 18), nb.empty(B), nb.empty(B, dtype=torch.uint8)
for rep in range(6):
    a = torch.as_tensor(rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6)), device="cuda")
    torch.cuda.synchronize()
    nb.step(a, None, out=(obs, rew, term), extras={"solver_stats": stats})
torch.cuda.synchronize()
ph = stats[4 * B:].view(B, 16).cpu().numpy().astype(np.int64)
nit = ph[:, 10]
names = ["constants staged + launch wait", "instance wait + prologue loads", "load / p_pot clamps", "projections + SoC",
         "bus injections", "Newton-Raphson", "branch flows + reward", "gather_full_state", "epilogue stores", "fence + publish"]
ok = (nit > 0) & (nit < 100)
print("B=%d, ordinary instances: %d, divergent: %d; Newton iterations mean %.2f" % (B, ok.sum(), (nit >= 100).sum(), nit[ok].mean()))
d = np.diff(np.concatenate([np.zeros((B, 1), dtype=np.int64), ph[:, :10]], axis=1), axis=1)
print("%-34s %9s %9s %9s" % ("phase (SM cycles)", "median", "p10", "p90"))
for k, n in enumerate(names):
    print("%-34s %9.0f %9.0f %9.0f" % (n, np.median(d[ok, k]), np.percentile(d[ok, k], 10), np.percentile(d[ok, k], 90)))
print("%-34s %9.0f %9.0f %9.0f" % ("total (kernel entry -> published)", np.median(ph[ok, 9]), np.percentile(ph[ok, 9], 10), np.percentile(ph[ok, 9], 90)))
