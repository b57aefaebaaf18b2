"""GPU tests of the batched LP solver (include/anm_lp.h) and of the MPC agents' device path.  The file sorts last on
purpose: the step path's parity tests run first.  Checker: the host build of the same solver code and SciPy's HiGHS on
the full DC-OPF (objective values -- LP optima are not unique); the stepped batch is replayed through the C oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bounds_from_tables(spec, agent, red, B, rng, prev=None):
    """Bounds of B programs: a random time of day / curtailment / state of charge per instance, or -- given the
    previous draw -- the next quarter of an hour with slightly different values (what consecutive MPC steps see)."""
    from gym_anm_b200 import lp as LP

    if prev is None:
        t = rng.integers(0, spec.table.shape[0], size=B)
        scale = rng.uniform(0.0, 1.0, size=(B, agent.n_gen, 1))
        soc = rng.uniform(agent.soc_min, agent.soc_max, size=(B, agent.n_des))
    else:
        t, scale, soc = prev
        t = (t + 1) % spec.table.shape[0]
        scale = np.clip(scale + rng.normal(0.0, 0.02, size=scale.shape), 0.0, 1.0)
        soc = np.clip(soc + rng.normal(0.0, 0.01, size=soc.shape), agent.soc_min, agent.soc_max)
    rows = spec.table[t] / agent.baseMVA  # [B, n_load + n_gen]: loads (<= 0), generation potentials
    N = agent.planning_steps
    Lf = np.repeat(rows[:, : agent.n_load, None], N, axis=2)
    Gf = np.repeat(rows[:, agent.n_load:, None], N, axis=2) * scale
    return LP.instance_bounds(red, agent, Lf, Gf, soc) + ((t, scale, soc),)


@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_device_solver_equals_host_build_cold_warm_and_restart(monkeypatch, kernel):
    """Both device kernels (one thread / one warp per program; ANM_LP_KERNEL) against the host build of their own code
    and against HiGHS."""
    import torch

    monkeypatch.setenv("ANM_LP_KERNEL", kernel)

    import test_mpc_agent as tm
    from gym_anm_b200 import lp as LP
    from gym_anm_b200.agents import MPCAgentConstant
    from gym_anm_b200.env_spec import anm6easy_spec

    spec = anm6easy_spec()
    env = tm._Env(spec, 1)
    agent = MPCAgentConstant(env.simulator, env.action_space, 0.995, safety_margin=0.96, planning_steps=10)
    red = LP.reduce_dcopf(agent)
    B = 1000  # not a multiple of the warp size: stride 1024
    dev = LP.BatchedLP(red.A, red.c, B, "cuda:0")
    assert dev.stride == 1024 and dev.bytes > red.m * red.n * 8 * B and dev.kernel == kernel
    rng = np.random.default_rng(7)
    state = prev = None
    for rnd in range(4):
        lo, up, prev = _bounds_from_tables(spec, agent, red, B, rng, prev)
        restart = (rng.random(B) < 0.3).astype(np.uint8) if rnd == 2 else None
        dev.lo[:, :B] = torch.as_tensor(lo.T, device="cuda:0")
        dev.up[:, :B] = torch.as_tensor(up.T, device="cuda:0")
        dev.solve(None if restart is None else torch.as_tensor(restart, device="cuda:0"))
        x_h, obj_h, st_h, it_h, state = LP.solve_host(red, lo, up, state=state, restart=restart, kernel=kernel)
        torch.cuda.synchronize()
        st_d, obj_d, it_d = dev.status.cpu().numpy(), dev.obj.cpu().numpy(), dev.iters.cpu().numpy()
        assert (st_h == 0).all() and (st_d == 0).all()
        np.testing.assert_allclose(obj_d, obj_h, rtol=1e-9, atol=1e-9)
        assert float(dev.violation().max()) <= 1e-8
        x_d = dev.x[:, :B].t().cpu().numpy()
        np.testing.assert_allclose(x_d @ red.c, obj_d, rtol=1e-12, atol=1e-12)
        if rnd == 0:
            assert it_d.min() > 0
        else:  # warm starts: a fraction of the cold pivots (restarted instances aside)
            keep = np.ones(B, bool) if restart is None else restart == 0
            assert it_d[keep].mean() < 0.6 * it0
        it0 = it_d.mean() if rnd == 0 else it0
        # same code, same inputs: the pivot sequences agree unless a rounding difference (FMA contraction) flips a tie
        assert (it_d == it_h).mean() > 0.9
        for i in rng.choice(B, size=4, replace=False):  # and HiGHS on the full program for a few
            Lf = lo[i, red.col_load].reshape(agent.n_load, -1)
            Gf = up[i, red.col_gen].reshape(agent.n_gen, -1)
            soc = agent.soc_max - up[i, red.n + red.row_soc[:, 0]]
            res = agent.solve_one(Lf, Gf, soc)[1]
            assert abs(res.fun - obj_d[i]) <= 1e-8 * max(1.0, abs(res.fun))
    dev.close()


@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_mpc_agents_on_device_drive_the_batched_env(monkeypatch, kernel):
    """BASELINE config 5 with the whole loop on the GPU: state tensor -> bounds -> batched LP -> action tensor -> step.
    Objective values against HiGHS on the full program (a sample), the stepped batch against the C oracle."""
    import torch

    monkeypatch.setenv("ANM_LP_KERNEL", kernel)

    import anm_oracle
    from golden_util import rel_err
    from gym_anm_b200.agents import MPCAgentConstant, MPCAgentPerfect
    from gym_anm_b200.anm6 import BatchedANM6Easy

    for cls, N in ((MPCAgentConstant, 10), (MPCAgentPerfect, 6)):
        B = 256
        env = BatchedANM6Easy(B, validate_actions=True)
        env.reset(seed=16384)
        agent = cls(env.simulator, env.action_space, env.gamma, safety_margin=0.96, planning_steps=N, device="cuda:0",
                    refresh=4)
        host = cls(env.simulator, env.action_space, env.gamma, safety_margin=0.96, planning_steps=N)
        cpu = anm_oracle.OracleEnv(env.spec, B)
        soc, aux, term = env.native.get_state()
        cpu.soc[:], cpu.aux[:], cpu.terminated[:] = soc.cpu().numpy(), aux.cpu().numpy(), 0
        total = 0.0
        for t in range(6):
            a = agent.act(env)
            assert isinstance(a, torch.Tensor) and a.is_cuda and a.shape == (B, 6)
            st = env.state.cpu().numpy()
            p_load, p_gen_max, s = host.state_to_pu(st)
            Lf, Gf = host.forecast_batch(env, p_load, p_gen_max)
            obj = agent._dev.lp.obj.cpu().numpy()
            for i in range(0, B, 37):
                res = host.solve_one(Lf[i], Gf[i], s[i])[1]
                assert abs(res.fun - obj[i]) <= 1e-8 * max(1.0, abs(res.fun))
            obs_g, r_g, term_g, _, _ = env.step(a)
            obs_c, r_c, term_c, _ = cpu.step(a.cpu().numpy())
            assert not term_c.any() and np.array_equal(term_g.cpu().numpy(), term_c)
            assert rel_err(obs_g.cpu().numpy(), obs_c) < 1e-8 and rel_err(r_g.cpu().numpy(), r_c) < 1e-8
            total += float(r_g.mean())
        assert total / 6 > -5.0  # the MPC policy operates the grid at low cost
        assert agent.lp_stats["host_fallbacks"] == 0 and agent.lp_stats["solves"] == 6
        assert agent._dev.lp.kernel == kernel
        agent.close()
