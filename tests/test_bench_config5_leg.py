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
