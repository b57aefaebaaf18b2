"""Smallest run through the parking path (debug aid): a few instances, a few steps, compared with ANM_PARK=0 results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_anm_b200.anm6 import BatchedANM6Easy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
env = BatchedANM6Easy(B, validate_actions=False)
print("sizes", env.native.sizes, flush=True)
env.reset(seed=2020)
torch.cuda.synchronize()
print("reset ok", flush=True)
rng = np.random.default_rng(0)
acts = rng.uniform(env.spec.action_low, env.spec.action_high, size=(T, B, 6))
tot = 0
for t in range(T):
    o, r, d, _, _ = env.step(acts[t])
    torch.cuda.synchronize()
    tot += int(d.sum())
print("steps ok, terminated", tot, "checksum", float(env._obs.sum()), flush=True)
env.native.set_autoreset_pool(env.state.clone())
obs, rew, term = env.native.rollout(torch.as_tensor(acts, device=env.device))
torch.cuda.synchronize()
print("rollout ok", float(obs.sum()), int(term.sum()), flush=True)
