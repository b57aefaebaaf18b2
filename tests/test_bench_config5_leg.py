"""bench.py's closed-loop MPC leg (BASELINE configs[4]) driven on the CPU: the batched environment is replaced by a
stand-in over the C oracle, the device LP handle by its host stand-in and torch.cuda's timing calls by host clocks, so
that the leg's own logic (sharding, estimate -> step count, timed loop, counters, oracle replay) runs without a GPU."""
import time
import types

import numpy as np
import torch

import bench
import test_mpc_agent as tm
from gym_anm_b200 import agents
from gym_anm_b200 import lp as LP
from gym_anm_b200.env_spec import anm6easy_spec


class _FakeNative:
    def __init__(self, env):
        self.env, self.launch_count = env, 0

    def get_state(self):
        c = self.env._e.cpu
        return (torch.as_tensor(c.soc.copy()), torch.as_tensor(c.aux.copy()),
                torch.as_tensor(np.asarray(c.terminated).astype(np.uint8)))


class _FakeBatchedEnv:
    def __init__(self, B, device=None, env_offset=0, validate_actions=False):
        self.spec = anm6easy_spec()
        self._e = tm._Env(self.spec, B, seed=2020 + env_offset)
        self.simulator, self.action_space, self.gamma = self._e.simulator, self._e.action_space, 0.995
        self.P_loads, self.P_maxs = self._e.P_loads, self._e.P_maxs
        self.native = _FakeNative(self)
        self._term = np.zeros(B, bool)

    def reset(self, seed=None):
        return None, {}

    @property
    def state(self):
        return torch.as_tensor(self._e.state)

    @property
    def terminated(self):
        return torch.as_tensor(self._term)

    def step(self, act):
        obs, r, term = self._e.step(act.numpy())
        self._term = np.asarray(term).astype(bool)
        self.native.launch_count += 1
        return torch.as_tensor(obs), torch.as_tensor(r), torch.as_tensor(self._term), None, {}


class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1000.0 * (other.t - self.t)


def test_config5_leg_logic_on_cpu(monkeypatch):
    from lp_host_standin import HostWarpLP

    import gym_anm_b200.anm6 as anm6

    monkeypatch.setattr(anm6, "BatchedANM6Easy", _FakeBatchedEnv)
    monkeypatch.setattr(LP, "BatchedLP", HostWarpLP)
    monkeypatch.setattr(agents.MPCAgentConstant, "act", agents.MPCAgentConstant.act_device)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    legs = [bench.Config5Leg(40, 2, rank, torch.device("cpu"), n_check=8) for rank in (0, 1)]
    assert [leg.B for leg in legs] == [20, 20]
    est = max(leg.estimate() for leg in legs)
    assert est > 0 and len(legs[0].rec) == 8
    n = 4
    for leg in legs:
        ms = leg.timed(n)
        assert ms > 0
        o = leg.out
        assert o["B"] == 20 and o["lp_kernel"] == "warp" and o["launches"] == 2 * (8 + n)
        assert o["lp_stats"]["solves"] == 8 + n and o["lp_stats"]["host_fallbacks"] == 0
        assert 0.0 <= o["mean_pivots_per_solve"] < 20.0 and o["terminated_frac"] == 0.0
        chk = leg.check()
        assert chk == {"instances": 8, "steps": 8, "max_rel_err_obs": 0.0, "terminated_equal": True}
        leg.close()
    # the two shards start from different instances (global index = seed offset)
    assert not np.array_equal(legs[0].chk0[0], legs[1].chk0[0]) or not np.array_equal(legs[0].chk0[1], legs[1].chk0[1])
    assert isinstance(types.SimpleNamespace(), object)


class _StubLeg:
    """A leg whose stages succeed or fail on demand (which rank, which stage)."""

    fail = {}  # stage -> set of ranks

    def __init__(self, envs_global, world, rank, dev):
        self.rank, self.closed = rank, False
        if rank in self.fail.get("init", ()):
            raise RuntimeError("init failed on rank %d" % rank)

    def estimate(self):
        if self.rank in self.fail.get("estimate", ()):
            raise RuntimeError("estimate failed on rank %d" % self.rank)
        return 0.01 * (1 + self.rank)

    def timed(self, n):
        if self.rank in self.fail.get("timed", ()):
            raise RuntimeError("timed failed on rank %d" % self.rank)
        self.out = {"B": 8, "mean_pivots_per_solve": 2.0, "lp_kernel": "warp", "lp_bytes": 1, "lp_stats": {"solves": n},
                    "mean_reward": -1.0, "terminated_frac": 0.0, "launches": 2 * n}
        return 10.0 * n * (1 + self.rank)

    def check(self):
        return {"instances": 1, "steps": 1, "max_rel_err_obs": 0.0, "terminated_equal": True}

    def close(self):
        self.closed = True


def test_run_config5_branches_single_process():
    calls = {"barrier": 0}

    def barrier():
        calls["barrier"] += 1

    maxr = lambda v: list(v)  # noqa: E731
    _StubLeg.fail = {}
    ok = bench.run_config5(16, 0.35, 1, 0, None, barrier, maxr, leg_factory=_StubLeg)
    assert ok["steps"] == 35 and abs(ok["ms_per_step"] - 10.0) < 1e-12 and abs(ok["value"] - 16 / 0.010) < 1e-6
    assert ok["oracle_check"]["terminated_equal"] and ok["lp_kernel"] == "warp" and calls["barrier"] == 2
    for stage in ("init", "estimate", "timed"):
        _StubLeg.fail = {stage: {0}}
        bad = bench.run_config5(16, 0.35, 1, 0, None, barrier, maxr, leg_factory=_StubLeg)
        assert set(bad) == {"error"} and stage in bad["error"]
    _StubLeg.fail = {}


_C5_WORKER = r"""
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests"), os.path.join({root!r}, "oracle")]
import json
import torch, torch.distributed as dist
rank = int(sys.argv[1]); stage = sys.argv[2]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="{port}")
dist.init_process_group("gloo", rank=rank, world_size=2)
import bench
from test_bench_config5_leg import _StubLeg
_StubLeg.fail = {{stage: {{1}}}} if stage != "none" else {{}}
def maxr(vals):
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
out = bench.run_config5(16, 0.35, 2, rank, None, dist.barrier, maxr, leg_factory=_StubLeg)
print("C5", rank, json.dumps(out))
dist.barrier(); dist.destroy_process_group()
"""


def test_run_config5_world2_a_failing_rank_does_not_hang_the_other(tmp_path):
    """N > 1 (gloo, two processes): rank 1 fails in each stage in turn; both ranks return (no rank waits in a
    collective for one that has given up), both report the error; with no failure the time is the max over ranks."""
    import json
    import os
    import socket
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for stage in ("none", "init", "estimate", "timed"):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        script = tmp_path / ("w_%s.py" % stage)
        script.write_text(_C5_WORKER.format(root=root, port=port))
        procs = [subprocess.Popen([sys.executable, str(script), str(r), stage], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT) for r in range(2)]
        outs = [p.communicate(timeout=240)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        res = [json.loads([ln for ln in o.splitlines() if ln.startswith("C5 ")][0].split(" ", 2)[2]) for o in outs]
        if stage == "none":
            # est = max(0.01, 0.02) -> 18 steps; rank 1 is the slow one: 20 ms per step
            assert all(r["steps"] == 18 and abs(r["ms_per_step"] - 20.0) < 1e-9 for r in res), res
        else:
            assert all(set(r) == {"error"} for r in res), res
            assert stage in res[1]["error"] and (stage in res[0]["error"] or "another rank" in res[0]["error"])


class _StubLeg4:
    def __init__(self, B, world, rank, dev):
        self.B = B

    def estimate(self):
        return 0.02

    def timed(self, n):
        self.out = {"launches": n, "workload": "stub feeder", "terminated_frac_last_step": 0.0, "lanes_per_env": 32}
        return 20.0 * n

    def check(self):
        return None

    def close(self):
        pass


def test_run_config4_line():
    out = bench.run_config4(8192, 0.35, 2, 0, None, lambda: None, lambda v: list(v), leg_factory=_StubLeg4)
    # 18 blocks of 20 steps, 20 ms per block: 1 ms per step for 2 x 8192 instances
    assert out["steps"] == 360 and out["steps_per_launch"] == 20 and abs(out["ms_per_step"] - 1.0) < 1e-12
    assert abs(out["value"] - 2 * 8192 / 1e-3) < 1e-3 and out["gpu_launches"] == 18 and "stub feeder" in out["what"]

    class Broken(_StubLeg4):
        def timed(self, n):
            raise MemoryError("out of memory")

    bad = bench.run_config4(8192, 0.35, 1, 0, None, lambda: None, lambda v: list(v), leg_factory=Broken)
    assert bad == {"error": "MemoryError: out of memory"}


class _FakeNB:
    O = 5
    sizes = {"lanes_per_env": 32}

    def __init__(self, B):
        self.B, self.launch_count, self.calls = B, 0, []

    def empty(self, *shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype)

    def rollout(self, actions, next_vars, out, chained):
        assert actions.shape[0] == next_vars.shape[0] == out[0].shape[0] == bench.Config4Leg.T
        assert actions.is_contiguous() and next_vars.is_contiguous()
        self.calls.append((int(actions[0, 0, 0]), chained))  # ring slot of the first step (the fake ring holds its index)
        out[2][:] = 0
        self.launch_count += 1


def test_config4_leg_indexing_on_cpu(monkeypatch):
    B, R = 6, 64

    def fake_setup(Bx, dev, rank):
        if rank == 1:
            raise AssertionError("an initial state without a solution")
        ring = torch.arange(R, dtype=torch.float64)[:, None, None].expand(R, Bx, 3).contiguous()
        return {"nb": _FakeNB(Bx), "ring": ring, "ring_nv": ring.clone(), "workload": "fake feeder"}

    monkeypatch.setattr(bench, "setup_config4", fake_setup)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    leg = bench.Config4Leg(B, 2, 1, torch.device("cpu"))  # rank 1's draw fails: rank 0's is used
    assert leg.seed_rank == 0
    assert leg.estimate() >= 0.0
    nb = leg.nb
    assert [c[1] for c in nb.calls] == [False, False, True, True]  # untimed block, then three with the 2nd / 3rd chained
    ms = leg.timed(5)
    assert ms >= 0.0 and leg.out["launches"] == 5 and leg.out["seed_rank"] == 0 and leg.out["workload"] == "fake feeder"
    starts = [c[0] for c in nb.calls]
    assert starts == [(b * 20) % (R - 20 + 1) for b in range(9)] and max(starts) + 20 <= R
    assert [c[1] for c in nb.calls[4:]] == [False, True, True, True, True]
    leg.close()
