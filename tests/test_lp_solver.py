"""The batched LP solver behind the MPC action source (include/anm_lp.h, gym_anm_b200/lp.py), WITHOUT a GPU: the
solver's code is compiled for the host too (anm_debug_lp_solve_host) and checked here against SciPy's HiGHS --
random boxed programs, the reduced DC-OPF of the MPC agents in closed loop (warm starts), the reference test's own
constraints C1..C7 (tests/test_dcopf_agent.py:64-107) on the reconstructed full solution, and the agents' device
path end to end through a host stand-in of the device handle (tests/lp_host_standin.py)."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
from scipy.optimize import linprog

import test_mpc_agent as tm
from gym_anm_b200 import _capi
from gym_anm_b200 import lp as LP
from gym_anm_b200.agents import MPCAgentConstant, MPCAgentPerfect
from gym_anm_b200.env_spec import anm6easy_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lp_header_symbols_exported_and_errors_are_loud():
    hdr = open(os.path.join(ROOT, "include", "anm_lp.h")).read()
    declared = set(re.findall(r"\b(anm_(?:debug_)?lp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.LP_EXPORTED_SYMBOLS)
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    A, c = np.eye(2), np.ones(2)
    h = ctypes.c_void_p()
    dp = lambda a: a.ctypes.data_as(_capi.c_double_p)  # noqa: E731
    assert lib.anm_lp_create(2, 2, dp(A), dp(c), 0, 0, 0, 0, ctypes.byref(h)) == -1  # empty batch
    assert b"batch" in lib.anm_lp_last_error()
    assert lib.anm_lp_create(0, 2, dp(A), dp(c), 4, 4, 0, 0, ctypes.byref(h)) == -1
    import torch

    if not torch.cuda.is_available():  # no device: creation fails (no CPU fallback), and says why
        assert lib.anm_lp_create(2, 2, dp(A), dp(c), 4, 32, 0, 0, ctypes.byref(h)) == -2
        assert b"no CPU fallback" in lib.anm_lp_last_error()
        with pytest.raises(Exception, match="CUDA device is required"):
            LP.BatchedLP(A, c, 4, "cuda:0")


def _random_program(rng, n, m, density):
    A = rng.normal(size=(m, n)) * (rng.random((m, n)) < density)
    c = np.round(rng.normal(size=n), 2) * (rng.random(n) < 0.8)
    return np.ascontiguousarray(A), np.ascontiguousarray(c)


def _random_bounds(rng, A, B, width=1.0):
    """Feasible by construction: boxes around a random point and around its row activities."""
    m, n = A.shape
    x0 = rng.normal(size=(B, n))
    lo_c, up_c = x0 - rng.random((B, n)) * 2, x0 + rng.random((B, n)) * 2
    fixed = rng.random((B, n)) < 0.1
    lo_c[fixed] = up_c[fixed] = x0[fixed]
    act = x0 @ A.T
    lo_r, up_r = act - rng.random((B, m)) * width, act + rng.random((B, m)) * width
    lo_r[rng.random((B, m)) < 0.4] = -np.inf  # one-sided rows
    up_r[rng.random((B, m)) < 0.2] = np.inf
    eq = rng.random((B, m)) < 0.1
    lo_r[eq] = up_r[eq] = act[eq]
    return np.concatenate([lo_c, lo_r], axis=1), np.concatenate([up_c, up_r], axis=1)


def _highs(A, c, lo, up):
    m, n = A.shape
    A_ub, b_ub = [], []
    for i in range(m):
        if np.isfinite(up[n + i]):
            A_ub.append(A[i]), b_ub.append(up[n + i])
        if np.isfinite(lo[n + i]):
            A_ub.append(-A[i]), b_ub.append(-lo[n + i])
    return linprog(c, A_ub=np.array(A_ub) if A_ub else None, b_ub=np.array(b_ub) if b_ub else None,
                   bounds=np.stack([lo[:n], up[:n]], axis=1), method="highs")


@pytest.mark.parametrize("kernel", ["thread", "warp"])
@pytest.mark.parametrize("n,m,density", [(6, 4, 1.0), (12, 20, 0.5), (30, 25, 0.2), (40, 60, 0.15), (70, 33, 0.3),
                                         (16, 300, 0.2)])  # the last: more rows than the warp kernel keeps in its scratch
def test_random_boxed_programs_match_highs_cold_and_warm(n, m, density, kernel):
    rng = np.random.default_rng(100 * n + m)
    A, c = _random_program(rng, n, m, density)
    red = types.SimpleNamespace(A=A, c=c, n=n, m=m)
    B, state, state_rev = 24, None, None
    for rnd in range(4):  # round 0 cold, then new bounds on the old bases; round 3 restarts every other instance
        lo, up = _random_bounds(rng, A, B)
        restart = (np.arange(B) % 2).astype(np.uint8) if rnd == 3 else None
        x, obj, status, iters, state = LP.solve_host(red, lo, up, state=state, restart=restart, kernel=kernel)
        if kernel == "warp":  # the lanes of every phase in reverse order: bit-identical, tableaux included
            x2, obj2, status2, iters2, state_rev = LP.solve_host(red, lo, up, state=state_rev, restart=restart,
                                                                 kernel="warp", reverse_lanes=True)
            assert np.array_equal(x, x2) and np.array_equal(obj, obj2) and np.array_equal(iters, iters2)
            assert np.array_equal(status, status2) and np.array_equal(state, state_rev)
        assert (status == 0).all(), status
        act = np.concatenate([x, x @ A.T], axis=1)
        assert np.maximum(lo - act, act - up).max() <= 1e-8
        for i in range(B):
            ref = _highs(A, c, lo[i], up[i])
            assert ref.status == 0
            assert abs(ref.fun - obj[i]) <= 1e-8 * max(1.0, abs(ref.fun)), (rnd, i, ref.fun, obj[i])
        np.testing.assert_allclose(obj, x @ c, rtol=1e-12, atol=1e-12)


def test_infeasible_and_unboxed_programs_are_reported():
    A = np.array([[1.0, 1.0], [1.0, -1.0]])
    c = np.array([1.0, 1.0])
    red = types.SimpleNamespace(A=A, c=c, n=2, m=2)
    inf = np.inf
    # instance 0: x1 + x2 >= 5 inside the box [0, 1]^2: infeasible; instance 1: feasible; instance 2: a cost that pulls
    # towards an infinite bound
    lo = np.array([[0.0, 0.0, 5.0, -inf], [0.0, 0.0, 1.0, -inf], [-inf, 0.0, 1.0, -inf]])
    up = np.array([[1.0, 1.0, inf, inf], [1.0, 1.0, inf, inf], [1.0, 1.0, inf, inf]])
    for kernel in ("thread", "warp"):
        x, obj, status, iters, _ = LP.solve_host(red, lo, up, kernel=kernel)
        assert list(status) == [1, 0, 3]
        assert abs(obj[1] - 1.0) < 1e-12
    # an iteration cap of one pivot on a program that needs more
    rng = np.random.default_rng(5)
    A, c = _random_program(rng, 12, 20, 0.5)
    lo, up = _random_bounds(rng, A, 8)
    red = types.SimpleNamespace(A=A, c=c, n=12, m=20)
    for kernel in ("thread", "warp"):
        _, _, status, iters, _ = LP.solve_host(red, lo, up, max_iter=1, kernel=kernel)
        assert set(status) <= {0, 2} and (status == 2).any() and iters.max() == 1


def _full_solution(red, x):
    xf = np.zeros(red.n_full)
    xf[red.keep] = x
    xf[red.elim] = red.G @ x
    return xf


@pytest.mark.parametrize("N,cls", [(1, MPCAgentConstant), (3, MPCAgentConstant), (10, MPCAgentConstant),
                                   (5, MPCAgentPerfect)])
def test_reduced_dcopf_closed_loop_matches_highs(N, cls):
    """The agents' program in the solver's form: same optimal value as HiGHS on the full program at every step of a
    closed loop driven by the solver's own actions (warm starts from step 1 on), and the reference test's constraints
    on the full solution (angles, slack and storage injections reconstructed from the kept columns)."""
    spec = anm6easy_spec()
    B = 24
    env = tm._Env(spec, B, seed=3)
    agent = cls(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=N)
    red = LP.reduce_dcopf(agent)
    assert red.n == N * (agent.n_load + agent.n_gen + 2 * agent.n_des + agent.n_branch)
    assert red.m == N * (2 * agent.n_branch + 2 * agent.n_des)
    state, state_w, state_wr, pivots = None, None, None, []
    for t in range(6):
        p_load, p_gen_max, soc = agent.state_to_pu(env.state)
        Lf, Gf = agent.forecast_batch(env, p_load, p_gen_max)
        lo, up = LP.instance_bounds(red, agent, Lf, Gf, soc)
        x, obj, status, iters, state = LP.solve_host(red, lo, up, state=state)
        assert (status == 0).all()
        # the warp kernel's code (lane phases as loops, forwards and backwards): same optimum, same pivots
        xw, objw, stw, itw, state_w = LP.solve_host(red, lo, up, state=state_w, kernel="warp")
        xr, objr, str_, itr, state_wr = LP.solve_host(red, lo, up, state=state_wr, kernel="warp", reverse_lanes=True)
        assert np.array_equal(xw, xr) and np.array_equal(objw, objr) and np.array_equal(state_w, state_wr)
        assert (stw == 0).all() and (itw == iters).mean() > 0.8
        np.testing.assert_allclose(objw, obj, rtol=1e-10, atol=1e-10)
        actw = np.concatenate([xw, xw @ red.A.T], axis=1)
        assert np.maximum(lo - actw, actw - up).max() <= 1e-9
        pivots.append(iters.mean())
        assert not (np.abs(x[:, red.free_cols]) > LP.BIG / 2).any()
        assert np.abs(x @ red.theta_map.T).max() < np.pi
        for i in range(B):
            act_ref, res = agent.solve_one(Lf[i], Gf[i], soc[i])
            assert abs(res.fun - obj[i]) <= 1e-9 * max(1.0, abs(res.fun))
            xf = _full_solution(red, x[i])
            np.testing.assert_allclose(agent.c @ xf, obj[i], rtol=1e-10, atol=1e-10)
            if cls is MPCAgentConstant:
                tm._check_constraints(agent, spec, env.state[i], types.SimpleNamespace(x=xf), N)
        P_gen = x[:, red.col_gen[:, 0]] * agent.baseMVA
        P_des = (x @ red.A[red.row_pdes[:, 0]].T) * agent.baseMVA
        a = np.concatenate([P_gen, np.zeros_like(P_gen), P_des, np.zeros_like(P_des)], axis=1)
        env.step(np.clip(a, env.action_space.low, env.action_space.high))
    assert max(pivots[1:]) < 0.5 * pivots[0] + 1.0, pivots  # the warm start is worth something


class _TensorEnv:  # what `act_device` reads from a batched environment
    def __init__(self, env):
        import torch

        self.state = torch.as_tensor(env.state)
        self.P_loads, self.P_maxs = env.P_loads, env.P_maxs


@pytest.mark.parametrize("cls,N,kernel", [(MPCAgentConstant, 10, "thread"), (MPCAgentPerfect, 4, "thread"),
                                          (MPCAgentConstant, 10, "warp")])
def test_agent_device_path_through_host_standin(monkeypatch, cls, N, kernel):
    """`MPCAgent(device=...).act(env)` end to end (torch bounds scatter, checks, action extraction, periodic refresh,
    second solve and host fallback) with the device handle replaced by its host stand-in."""
    from lp_host_standin import HostLP, HostWarpLP

    monkeypatch.setattr(LP, "BatchedLP", HostLP if kernel == "thread" else HostWarpLP)
    spec = anm6easy_spec()
    B = 20
    env = tm._Env(spec, B, seed=11)
    dev = cls(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=N, device="cpu", refresh=3)
    host = cls(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=N)
    for t in range(7):
        tenv = _TensorEnv(env)
        a = dev.act_device(tenv)
        assert a.shape == (B, 6) and a.dtype.is_floating_point
        a = a.numpy()
        assert (a >= env.action_space.low - 1e-12).all() and (a <= env.action_space.high + 1e-12).all()
        p_load, p_gen_max, soc = host.state_to_pu(env.state)
        Lf, Gf = host.forecast_batch(env, p_load, p_gen_max)
        objs = np.array([host.solve_one(Lf[i], Gf[i], soc[i])[1].fun for i in range(B)])
        np.testing.assert_allclose(dev._dev.lp.obj.numpy(), objs, rtol=1e-9, atol=1e-9)
        if t in (3, 6):  # a refresh solve starts every instance from the all-slack basis again
            assert dev._dev.lp.iters.float().mean() > 5
        env.step(a)
    assert dev.lp_stats == {"solves": 7, "second_solves": 0, "host_fallbacks": 0}
    # instances flagged as failed: solved again from scratch; still flagged: the host LP supplies their actions
    calls = {"n": 0}
    real_failed = type(dev._dev).failed

    def failed(self):
        import torch

        calls["n"] += 1
        bad = real_failed(self)
        bad[1] = True  # instance 1 "fails" both times
        if calls["n"] == 1:
            bad[2] = True
        return bad

    monkeypatch.setattr(type(dev._dev), "failed", failed)
    from gym_anm_b200.errors import LPSolverError

    with pytest.raises(LPSolverError, match="1 of 20 programs"):  # the default: no silent CPU path
        dev.act_device(_TensorEnv(env))
    assert dev.lp_stats["second_solves"] == 1 and dev.lp_stats["host_fallbacks"] == 0
    calls["n"] = 0
    dev.on_lp_failure = "host"  # opt-in
    a = dev.act_device(_TensorEnv(env)).numpy()
    assert dev.lp_stats["second_solves"] == 2 and dev.lp_stats["host_fallbacks"] == 1
    p_load, p_gen_max, soc = host.state_to_pu(env.state)
    Lf, Gf = host.forecast_batch(env, p_load, p_gen_max)
    np.testing.assert_allclose(a[1], host.solve_one(Lf[1], Gf[1], soc[1])[0], atol=1e-12)


@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_reduced_dcopf_on_a_meshed_30_bus_feeder(kernel):
    """The reduction and the solver on a general network (config 4's synthetic 30-bus feeder: 29 + 2 branches, so the
    eliminated angles are a genuine linear solve, 10 loads, 6 generators, 3 storage units): the reconstructed full
    solution satisfies the agent's FULL program (equalities, inequalities, every bound) and has HiGHS's optimal value;
    cold, then warm from perturbed states."""
    from gym_anm_b200.env_spec import HostEnvSpec
    from gym_anm_b200.networks import synth_feeder_network
    from gym_anm_b200.spaces import Box

    spec = HostEnvSpec(synth_feeder_network(), "state", 1, 0.25, 0.99, 100, np.array([[0, 95]]), (1, 100))
    cn = spec.cn
    N, B = 3, 10
    agent = MPCAgentConstant(tm._Sim(spec), Box(spec.action_low, spec.action_high), 0.99, safety_margin=0.9, planning_steps=N)
    red = LP.reduce_dcopf(agent)
    nl = agent.n_branch
    assert (agent.n_bus, agent.n_load, agent.n_gen, agent.n_des) == (30, 10, 6, 3) and nl == 31
    assert red.n == N * (10 + 6 + 2 * 3 + nl) and red.m == N * (2 * nl + 2 * 3)
    rng = np.random.default_rng(4)
    p_load = rng.uniform([cn.devices[i].p_min for i in cn.load_ids], 0.0, size=(B, 10))
    p_gen = rng.uniform(0.0, [cn.devices[i].p_max for i in cn.gen_ids], size=(B, 6))
    soc = rng.uniform(agent.soc_min, agent.soc_max, size=(B, 3))
    state = None
    for rnd in range(3):
        Lf, Gf = np.repeat(p_load[:, :, None], N, axis=2), np.repeat(p_gen[:, :, None], N, axis=2)
        lo, up = LP.instance_bounds(red, agent, Lf, Gf, soc)
        x, obj, status, iters, state = LP.solve_host(red, lo, up, state=state, kernel=kernel)
        assert (status == 0).all()
        assert np.abs(x @ red.theta_map.T).max() < np.pi and not (np.abs(x[:, red.free_cols]) > LP.BIG / 2).any()
        for i in range(B):
            xf = _full_solution(red, x[i])
            lb, ub, b_ub = agent.instance_arrays(Lf[i], Gf[i], soc[i])
            assert np.abs(agent.A_eq @ xf).max() < 1e-9
            assert (agent.A_ub @ xf - b_ub).max() < 1e-8
            assert (lb - xf).max() < 1e-8 and (xf - ub).max() < 1e-8
            res = agent.solve_one(Lf[i], Gf[i], soc[i])[1]
            assert res.status == 0 and abs(res.fun - obj[i]) <= 1e-8 * max(1.0, abs(res.fun))
        if rnd > 0:
            assert iters.mean() < 0.7 * cold
        cold = iters.mean() if rnd == 0 else cold
        p_load = np.clip(p_load * rng.uniform(0.97, 1.03, p_load.shape), [cn.devices[i].p_min for i in cn.load_ids], 0.0)
        p_gen = np.clip(p_gen * rng.uniform(0.97, 1.03, p_gen.shape), 0.0, [cn.devices[i].p_max for i in cn.gen_ids])
        soc = np.clip(soc + rng.normal(0, 0.01, soc.shape), agent.soc_min, agent.soc_max)


@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_solutions_do_not_depend_on_the_batch_an_instance_is_in(kernel):
    """Shard equivalence of the action source (config 5 shards the instances over the GPUs): a program's pivots and
    solution are bit-identical whether it is solved in the whole batch or in a shard of it, cold and warm."""
    spec = anm6easy_spec()
    B = 12
    env = tm._Env(spec, B, seed=21)
    agent = MPCAgentConstant(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=4)
    red = LP.reduce_dcopf(agent)
    whole, lo_s, hi_s = None, None, None
    for t in range(3):
        p_load, p_gen_max, soc = agent.state_to_pu(env.state)
        Lf, Gf = agent.forecast_batch(env, p_load, p_gen_max)
        lo, up = LP.instance_bounds(red, agent, Lf, Gf, soc)
        x, obj, status, iters, whole = LP.solve_host(red, lo, up, state=whole, kernel=kernel)
        x0, obj0, _, it0, lo_s = LP.solve_host(red, lo[:5], up[:5], state=lo_s, kernel=kernel)
        x1, obj1, _, it1, hi_s = LP.solve_host(red, lo[5:], up[5:], state=hi_s, kernel=kernel)
        assert np.array_equal(x, np.concatenate([x0, x1])) and np.array_equal(obj, np.concatenate([obj0, obj1]))
        assert np.array_equal(iters, np.concatenate([it0, it1]))
        P_gen = x[:, red.col_gen[:, 0]] * agent.baseMVA
        P_des = (x @ red.A[red.row_pdes[:, 0]].T) * agent.baseMVA
        a = np.concatenate([P_gen, np.zeros_like(P_gen), P_des, np.zeros_like(P_des)], axis=1)
        env.step(np.clip(a, env.action_space.low, env.action_space.high))


@pytest.mark.parametrize("N", [1, 3, 20])
def test_dcopf_1000_step_rollouts_like_the_reference_test(N):
    """The reference's own test of this path (tests/test_dcopf_agent.py:29-62): 1000-step closed-loop rollouts of
    MPCAgentConstant with planning horizons 1, 3 and 20 on ANM6Easy, seed 2020, `gamma` = 0.9 (the reference passes its
    safety margin into that slot, :31), DC-OPF constraints C1..C7 checked at every step -- here with the programs
    solved by the warp kernel's code (host build), warm-started over the whole rollout without a single rebuild."""
    spec = anm6easy_spec()
    B = 2
    env = tm._Env(spec, B, seed=2020)
    agent = MPCAgentConstant(env.simulator, env.action_space, 0.9, planning_steps=N)
    red = LP.reduce_dcopf(agent)
    state, pivots = None, 0
    steps = 1000
    for t in range(steps):
        p_load, p_gen_max, soc = agent.state_to_pu(env.state)
        Lf, Gf = agent.forecast_batch(env, p_load, p_gen_max)
        lo, up = LP.instance_bounds(red, agent, Lf, Gf, soc)
        x, obj, status, iters, state = LP.solve_host(red, lo, up, state=state, kernel="warp")
        assert (status == 0).all(), (t, status)
        pivots += int(iters.sum())
        for i in range(B):
            tm._check_constraints(agent, spec, env.state[i], types.SimpleNamespace(x=_full_solution(red, x[i])), N)
        if t % 100 == 0:  # the optimal value against HiGHS now and then
            for i in range(B):
                res = agent.solve_one(Lf[i], Gf[i], soc[i])[1]
                assert abs(res.fun - obj[i]) <= 1e-9 * max(1.0, abs(res.fun))
        P_gen = x[:, red.col_gen[:, 0]] * agent.baseMVA
        P_des = (x @ red.A[red.row_pdes[:, 0]].T) * agent.baseMVA
        a = np.concatenate([P_gen, np.zeros_like(P_gen), P_des, np.zeros_like(P_des)], axis=1)
        _, _, term = env.step(np.clip(a, env.action_space.low, env.action_space.high))
        assert not term.any()  # the MPC policy keeps the grid solvable
    assert pivots / (steps * B) < 0.2 * (red.n + red.m)  # warm starts: a small fraction of a cold solve per step
