"""MPC (DC-OPF) agents: the LP restated for SciPy/HiGHS satisfies the constraints the reference's own
test checks (tests/test_dcopf_agent.py:64-107, C1..C7), for planning horizons 1, 3 and 20; and a
config-5 style run: MPC-constant actions drive the environment (CPU oracle here, GPU in test_gpu_parity)."""
import numpy as np
import pytest

import anm_numpy
import anm_oracle
from gym_anm_b200.agents import MPCAgentConstant, MPCAgentPerfect
from gym_anm_b200.env_spec import anm6easy_spec
from gym_anm_b200.spaces import Box


class _Sim:  # the attributes MPCAgent reads from a simulator (mpc.py:53-119)
    def __init__(self, spec):
        cn = spec.cn
        self.baseMVA, self.lamb, self.delta_t = cn.baseMVA, cn.lamb, cn.delta_t
        self.buses, self.devices, self.branches, self.Y_bus = cn.buses, cn.devices, cn.branches, cn.Y_bus_dense


class _Env:
    def __init__(self, spec, B, seed=0):
        self.spec, self.B = spec, B
        self.simulator = _Sim(spec)
        self.action_space = Box(spec.action_low, spec.action_high)
        self.P_loads, self.P_maxs = spec.table.T[:3], spec.table.T[3:]
        self.cpu = anm_oracle.OracleEnv(spec, B)
        s0 = np.stack([anm_numpy.anm6easy_init_state(spec, np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed + i))))
                       for i in range(B)])  # fmt: skip
        _, self.state, conv = self.cpu.reset(s0)
        assert conv.all()

    def step(self, a):
        obs, r, term, info = self.cpu.step(a)
        self.state = info["state"]
        return obs, r, term


def _check_constraints(agent, spec, state_row, res, N):
    cn, x, S = spec.cn, res.x, agent.stride
    p_load, p_gen_max, soc = agent.state_to_pu(state_row)
    B = cn.Y_bus_dense.imag
    for i in range(N):
        th = x[i * S: i * S + 6]
        P = x[i * S + 6: i * S + 13]
        # C1 nodal DC power balance
        flows = [B[0, 1] * (th[0] - th[1]),
                 B[1, 0] * (th[1] - th[0]) + B[1, 2] * (th[1] - th[2]) + B[1, 3] * (th[1] - th[3]),
                 B[2, 1] * (th[2] - th[1]) + B[2, 4] * (th[2] - th[4]) + B[2, 5] * (th[2] - th[5]),
                 B[3, 1] * (th[3] - th[1]), B[4, 2] * (th[4] - th[2]), B[5, 2] * (th[5] - th[2])]
        P_bus = [P[0], 0.0, 0.0, P[1] + P[2], P[3] + P[4], P[5] + P[6]]
        np.testing.assert_allclose(flows, P_bus, atol=1e-5)
        # C2 loads follow the (constant) forecast, C3/C4 generator and storage limits, C6/C7 angles
        np.testing.assert_allclose(P[[1, 3, 5]], p_load[0], atol=1e-5)
        for k, d in enumerate((2, 4)):
            assert cn.devices[d].p_min - 1e-9 <= P[d] <= min(cn.devices[d].p_max, p_gen_max[0, k]) + 1e-9
        assert cn.devices[6].p_min - 1e-9 <= P[6] <= cn.devices[6].p_max + 1e-9
        assert np.all(np.abs(th) <= np.pi + 1e-9) and abs(th[0]) < 1e-9
    # C5 SoC window over the whole horizon (charging / discharging efficiencies as in mpc.py:287-297)
    s = soc[0, 0]
    for i in range(N):
        ch, dis = x[i * S + agent.o_ch], x[i * S + agent.o_dis]
        assert ch >= -1e-9 and dis >= -1e-9
        np.testing.assert_allclose(x[i * S + 6 + 6], dis - ch, atol=1e-7)
        s = s + ch * 0.25 * 0.9 - dis * 0.25 / 0.9
        assert -1e-7 <= s <= 1.0 + 1e-7


@pytest.mark.parametrize("N", [1, 3, 20])
def test_mpc_constant_constraints(N):
    spec = anm6easy_spec()
    env = _Env(spec, 4)
    agent = MPCAgentConstant(env.simulator, env.action_space, 0.995, safety_margin=0.9, planning_steps=N)
    for t in range(12 if N < 20 else 4):
        p_load, p_gen_max, soc = agent.state_to_pu(env.state)
        Lf, Gf = agent.forecast_batch(env, p_load, p_gen_max)
        acts = []
        for i in range(env.B):
            a, res = agent.solve_one(Lf[i], Gf[i], soc[i])
            assert res.status == 0
            _check_constraints(agent, spec, env.state[i], res, N)
            assert env.action_space.contains(a) and a[2] == 0 and a[3] == 0 and a[5] == 0
            acts.append(a)
        _, r, term = env.step(np.stack(acts))
        assert not term.any()  # MPC keeps the network feasible
    assert np.all(r > -5.0)    # ... and cheap (random agents average around -40)


def test_mpc_perfect_uses_future_profiles():
    spec = anm6easy_spec()
    env = _Env(spec, 2, seed=3)
    agent = MPCAgentPerfect(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=4)
    p_load, p_gen_max, soc = agent.state_to_pu(env.state)
    Lf, Gf = agent.forecast_batch(env, p_load, p_gen_max)
    for i in range(2):
        t0 = int(env.state[i, -1])
        for k in range(4):
            np.testing.assert_allclose(Lf[i][:, k] * 100, env.P_loads[:, (t0 + 1 + k) % 96])
            np.testing.assert_allclose(Gf[i][:, k] * 100, env.P_maxs[:, (t0 + 1 + k) % 96])
    a = agent.act(env)
    assert a.shape == (2, 6)


def test_mpc_worker_pool_equals_serial():
    """BASELINE config 5's action source at scale: the LPs of a batch spread over worker processes give the same
    actions as the in-process loop (same LP, same solver, row order kept)."""
    spec = anm6easy_spec()
    rng = np.random.default_rng(3)
    space = Box(spec.action_low, spec.action_high)
    serial = MPCAgentConstant(_Sim(spec), space, 0.995, safety_margin=0.96, planning_steps=4)
    pooled = MPCAgentConstant(_Sim(spec), space, 0.995, safety_margin=0.96, planning_steps=4, workers=2)
    B = 24
    Lf = np.repeat(rng.uniform(-0.2, 0.0, (B, 3, 1)), 4, axis=2)
    Gf = np.repeat(rng.uniform(0.0, 0.4, (B, 2, 1)), 4, axis=2)
    soc = rng.uniform(0.0, 1.0, (B, 1))
    try:
        a, b = serial.solve_batch(Lf, Gf, soc), pooled.solve_batch(Lf, Gf, soc)
    finally:
        pooled.close()
    assert a.shape == b.shape == (B, 6) and np.array_equal(a, b)
