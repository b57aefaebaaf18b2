#!/usr/bin/env python
"""BASELINE config 4 (not the headline): synthetic 30-bus feeder, 8192 instances on one B200, random actions
and caller-supplied next_vars (device-resident rings), generic shared-memory kernel (58 x 59 fp64 Jacobian
per instance).  Prints env-steps/s from a CUDA-graph replay, like bench.py's `value`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_anm_b200.env_spec import HostEnvSpec  # noqa: E402
from gym_anm_b200.native import NativeBatch  # noqa: E402
from gym_anm_b200.networks import synth_feeder_network  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
spec = HostEnvSpec(synth_feeder_network(), "state", 1, 0.25, 0.99, 100, np.array([[0, 95]]), (1, 100))
cn = spec.cn
nb = NativeBatch(spec, B)
print("sizes:", {k: nb.sizes[k] for k in ("n_bus", "n_dev", "n_branch", "n_action", "n_obs", "lanes_per_env", "envs_per_block", "smem_bytes")})
rng = np.random.default_rng(30)
D, ns, ng = cn.N_device, cn.N_des, cn.N_non_slack_gen
pos = {d: k for k, d in enumerate(cn.devices)}
s0 = np.zeros((B, spec.state_N))
for i in cn.load_ids:
    s0[:, pos[i]] = rng.uniform(cn.devices[i].p_min * 100, 0, B)
for k, i in enumerate(cn.gen_ids):
    s0[:, pos[i]] = rng.uniform(0, cn.devices[i].p_max * 100, B)
    s0[:, 2 * D + ns + k] = rng.uniform(0, cn.devices[i].p_max * 100, B)
for k, i in enumerate(cn.des_ids):
    s0[:, 2 * D + k] = rng.uniform(0, cn.devices[i].soc_max * 100, B)
obs, state, conv = nb.reset(s0)
assert bool(conv.all())
nb.set_autoreset_pool(state.clone())
R = 64
dev = nb.device
lo, hi = torch.as_tensor(spec.action_low, device=dev), torch.as_tensor(spec.action_high, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(4)
acts = torch.rand((R, B, len(spec.action_low)), dtype=torch.float64, device=dev, generator=g) * (hi - lo) + lo
nv_lo = torch.as_tensor([cn.devices[i].p_min * 100 for i in cn.load_ids] + [0.0] * ng + [0.0], device=dev)
nv_hi = torch.as_tensor([0.0] * cn.N_load + [cn.devices[i].p_max * 100 for i in cn.gen_ids] + [95.0], device=dev)
nvs = torch.rand((R, B, spec.n_next_vars), dtype=torch.float64, device=dev, generator=g) * (nv_hi - nv_lo) + nv_lo
o, r, t = nb.empty(B, nb.O), nb.empty(B), nb.empty(B, dtype=torch.uint8)
nit = torch.zeros(B, dtype=torch.int32, device=dev)
for k in range(5):
    nb.step(acts[k], nvs[k], out=(o, r, t), extras={"n_iter": nit})
torch.cuda.synchronize()
graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        for k in range(R):
            nb.step(acts[k], nvs[k], out=(o, r, t), extras={"n_iter": nit})
torch.cuda.current_stream().wait_stream(side)
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (5 * R)
n = nit.cpu().numpy()
print("synth30: B=%d  %.3f ms/step  %.3g env-steps/s   last step: %d terminated, n_iter hist %s" % (
    B, ms, B / ms * 1e3, int(t.sum()), np.bincount(np.minimum(n, 101), minlength=102)[[0, 1, 2, 3, 4, 5, 100]]))
# the same as one rollout launch (R steps per instance, carried state on chip)
ro = (nb.empty(R, B, nb.O), nb.empty(R, B), nb.empty(R, B, dtype=torch.uint8))
nb.rollout(acts, nvs, out=ro)
torch.cuda.synchronize()
e0.record()
for k in range(5):
    nb.rollout(acts, nvs, out=ro, chained=(k > 0))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (5 * R)
print("synth30 rollout: B=%d  %.3f ms/step  %.3g env-steps/s" % (B, ms, B / ms * 1e3))
