"""Diagnostic (ANM_DIAG build): which Newton iterations trip the singular-block guard (device printf)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_anm_b200.anm6 import BatchedANM6Easy
B = 4096
env = BatchedANM6Easy(B, validate_actions=False)
env.reset(seed=2020)
rng = np.random.default_rng(0)
stats = torch.zeros(B, 20, dtype=torch.int32, device=env.device)
nit = torch.zeros(B, dtype=torch.int32, device=env.device)
for t in range(3):
    a = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))
    o, r, d = env.native.step(a, None, extras={"solver_stats": stats, "n_iter": nit})
    torch.cuda.synchronize()
    s = stats.cpu().numpy(); n = nit.cpu().numpy()
    print("step", t, "terminated", int(d.sum()), "tripped", int((s[:, 0] > 0).sum()), "n_iter of tripped", np.bincount(np.minimum(n[s[:,0]>0],101))[-5:] if (s[:,0]>0).any() else None, flush=True)
