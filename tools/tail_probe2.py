#!/usr/bin/env python
"""Cycles per Newton iteration of divergent instances in different company (in-kernel clock64)."""
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
for t in range(30):
    soc, aux, term = env.native.get_state()
    a = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))
    obs, r, term2, _, _ = env.step(a)
    idx = np.flatnonzero((term2 & ~term.bool()).cpu().numpy())
    S.append(soc.cpu().numpy()[idx]), X.append(aux.cpu().numpy()[idx]), A.append(a[idx])
S, X, A = np.concatenate(S), np.concatenate(X), np.concatenate(A)
soc, aux, _ = env.native.get_state()
ok = (~env.terminated).cpu().numpy()
oS, oX = soc.cpu().numpy()[ok], aux.cpu().numpy()[ok]
oA = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))[ok]


def probe(label, soc, aux, act):
    n = len(soc)
    e = BatchedANM6Easy(n, validate_actions=False)
    stats_all = torch.zeros(n * 20, dtype=torch.int32, device="cuda")  # diagnostic builds: [n, 4] + [n, 16] phase stamps
    stats = stats_all[: 4 * n].view(n, 4)
    nit = torch.zeros(n, dtype=torch.int32, device="cuda")
    obs, rew, term = e.native.empty(n, 18), e.native.empty(n), e.native.empty(n, dtype=torch.uint8)
    z = torch.zeros(n, dtype=torch.uint8, device="cuda")
    soc, aux, act = (torch.as_tensor(v, device="cuda").contiguous() for v in (soc, aux, act))
    for rep in range(3):
        e.native.set_state(soc, aux, z)
        e.native.step(act, None, out=(obs, rew, term), extras={"solver_stats": stats_all, "n_iter": nit})
    torch.cuda.synchronize()
    st, ni = stats.cpu().numpy().astype(np.int64), nit.cpu().numpy()
    d = ni >= 100
    loop = st[:, 2] * 16
    print("%-46s n=%4d divergent=%3d | divergent loop cycles/iter: median %5.0f max %5.0f | pass cycles max %d"
          % (label, n, d.sum(), np.median(loop[d]) / 100 if d.any() else 0, loop[d].max() / 100 if d.any() else 0, st[:, 3].max()))


probe("2 real divergent (one warp)", S[:2], X[:2], A[:2])
probe("1 real divergent + 1 ordinary (one warp)", np.concatenate([S[:1], oS[:1]]), np.concatenate([X[:1], oX[:1]]), np.concatenate([A[:1], oA[:1]]))
probe("32 real divergent", S[:32], X[:32], A[:32])
probe("same divergent instance x32", np.repeat(S[:1], 32, 0), np.repeat(X[:1], 32, 0), np.repeat(A[:1], 32, 0))
mS, mX, mA = oS[:4096].copy(), oX[:4096].copy(), oA[:4096].copy()
pos = np.arange(0, 35) * 97
mS[pos], mX[pos], mA[pos] = S[:35], X[:35], A[:35]
probe("35 real divergent among ordinary", mS, mX, mA)
mS, mX, mA = oS[:64].copy(), oX[:64].copy(), oA[:64].copy()
mS[0], mX[0], mA[0] = S[0], X[0], A[0]
probe("1 real divergent among 63 ordinary", mS, mX, mA)
