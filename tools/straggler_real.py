#!/usr/bin/env python
"""How long do REAL divergent instances take when they run alone?  Collect (state, action) pairs of
instances that terminate in a random-agent rollout, replay only those in a small batch and time it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

B = 4096
env = BatchedANM6Easy(B, validate_actions=False)
env.reset(seed=2020)
rng = np.random.default_rng(0)
S, A, X = [], [], []
for t in range(40):
    soc, aux, term = env.native.get_state()
    a = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))
    obs, r, term2, _, _ = env.step(a)
    new = (term2 & ~term.bool()).cpu().numpy()
    idx = np.flatnonzero(new)
    S.append(soc.cpu().numpy()[idx]), X.append(aux.cpu().numpy()[idx]), A.append(a[idx])
    nit = env.n_iter.cpu().numpy()
    if t < 3:
        print("step", t, "new terminations", len(idx), "n_iter of those", np.bincount(nit[idx])[-3:], "max n_iter others", nit[~new].max())
S, X, A = np.concatenate(S), np.concatenate(X), np.concatenate(A)
print("collected", len(S), "divergent (state, action) pairs")


def time_batch(n, soc, aux, act, label):
    e = BatchedANM6Easy(n, validate_actions=False)
    stats = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
    e._extras["solver_stats"] = stats
    z = torch.zeros(n, dtype=torch.uint8, device="cuda")
    soc, aux, act = (torch.as_tensor(v[:n], device="cuda").contiguous() for v in (soc, aux, act))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for rep in range(6):
        e.native.set_state(soc, aux, z)
        torch.cuda.synchronize()
        ev0.record()
        e.step(act)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1) * 1000)
    nit = e.n_iter.cpu().numpy()
    st = stats.cpu().numpy()
    print("%s: n=%d  %.1f us per launch (min of 6)   n_iter: min %d max %d   terminated %d | fallback iterations: mean %.1f max %d; |theta|>1e5 iterations: mean %.1f max %d" % (
        label, n, min(times), nit.min(), nit.max(), int(e.terminated.sum()), st[:, 0].mean(), st[:, 0].max(), st[:, 1].mean(), st[:, 1].max()))
    return nit


for n in (2, 32, 256):
    time_batch(n, S, X, A, "real divergent instances alone")
# the same number of ordinary instances
soc, aux, _ = env.native.get_state()
ok = (~env.terminated).cpu().numpy()
a = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))
time_batch(32, soc.cpu().numpy()[ok], aux.cpu().numpy()[ok], a[ok], "ordinary instances")
# one divergent instance per warp pair, rest ordinary (mixed batch like the bench)
mixS, mixX, mixA = soc.cpu().numpy()[ok][:4096].copy(), aux.cpu().numpy()[ok][:4096].copy(), a[ok][:4096].copy()
k = min(len(S), 35)
pos = rng.choice(len(mixS), k, replace=False)
mixS[pos], mixX[pos], mixA[pos] = S[:k], X[:k], A[:k]
time_batch(len(mixS), mixS, mixX, mixA, "mixed batch (%d divergent)" % k)
