#!/usr/bin/env python
"""Where does a bench-like launch spend its time?  Graph-replays steps like bench.py but with the
solver_stats diagnostics on, then prints per-launch device time and the in-kernel cycle counters of
ordinary vs divergent instances (SM cycles; 1965 MHz under load)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402

B = 4096
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
env = BatchedANM6Easy(B, validate_actions=False)
nb = env.native
env.reset(seed=2020)
nb.set_autoreset_pool(env.state.clone())
lo = torch.as_tensor(env.spec.action_low, device="cuda")
hi = torch.as_tensor(env.spec.action_high, device="cuda")
mid, half = (lo + hi) / 2, (hi - lo) / 2
g = torch.Generator(device="cuda")
g.manual_seed(1)
ring = (torch.rand((400, B, 6), dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * half * scale + mid
obs, rew, term = nb.empty(B, 18), nb.empty(B), nb.empty(B, dtype=torch.uint8)
stats = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
nit = torch.zeros(B, dtype=torch.int32, device="cuda")
ex = {"solver_stats": stats, "n_iter": nit}
for t in range(20):
    nb.step(ring[t], None, out=(obs, rew, term), extras=ex)
torch.cuda.synchronize()
# timed: graph of 200 steps, replayed 10x
G = 200
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
acc = torch.zeros_like(stats)      # per-instance maxima over all steps
fbsum = torch.zeros(B, dtype=torch.int64, device="cuda")
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        for t in range(G):
            nb.step(ring[t], None, out=(obs, rew, term), extras=ex)
            torch.maximum(acc, stats, out=acc)
            fbsum += stats[:, 0]
torch.cuda.current_stream().wait_stream(side)
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    graph.replay()
e1.record()
torch.cuda.synchronize()
print("action scale %.2f: %.1f us per step (graph replay, stats on)" % (scale, e0.elapsed_time(e1) * 1000 / (10 * G)))
st, ni = stats.cpu().numpy().astype(np.int64), nit.cpu().numpy()
loop = st[:, 2] * 16
strag = ni >= 100
print("last step: %d divergent instances; n_iter hist of the rest: %s" % (strag.sum(), np.bincount(ni[~strag])))
if strag.any():
    print("  divergent: Newton-loop cycles  min %d  median %d  max %d  -> %.0f cycles / iteration (median); fallback its max %d; |theta|>1e5 its max %d"
          % (loop[strag].min(), np.median(loop[strag]), loop[strag].max(), np.median(loop[strag]) / 100, st[strag, 0].max(), st[strag, 1].max()))
print("  ordinary : Newton-loop cycles  median %d  max %d  -> %.0f cycles / iteration;  whole pass cycles median %d  max %d"
      % (np.median(loop[~strag]), loop[~strag].max(), np.median(loop[~strag] / np.maximum(ni[~strag], 1)), np.median(st[~strag, 3]), st[~strag, 3].max()))
if strag.any():
    print("  divergent: whole pass cycles median %d max %d" % (np.median(st[strag, 3]), st[strag, 3].max()))

a = acc.cpu().numpy().astype(np.int64)
print("  over all steps: max Newton-loop cycles %d (%.0f / iteration), max pass cycles %d (%.1f us @1.965 GHz), max fallback its in one solve %d, fallback its total %d"
      % (a[:, 2].max() * 16, a[:, 2].max() * 16 / 100, a[:, 3].max(), a[:, 3].max() / 1965.0, a[:, 0].max(), int(fbsum.sum())))
