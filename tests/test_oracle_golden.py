"""CPU tests (`-m "not gpu"`): the oracles against the committed reference goldens.

Pins BOTH restatements used as checkers:
  * oracle/anm_numpy.py  -- same NumPy/SciPy calls as the reference: bit-exact (1e-12 across CPUs);
  * oracle/anm_oracle.c  -- scalar C port with dense LU: same `terminated` flags and NR
    iteration counts, values within 1e-9.
and, when /root/reference is present (build container), the NumPy port against the LIVE reference.
"""
import numpy as np
import pytest

import anm_numpy
import anm_oracle
import ref_loader
from golden_util import load, rel_err, transition_spec
from gym_anm_b200.env_spec import anm6easy_spec

NETS = ["2bus", "3bus_loop", "3bus_xfmr", "3bus_reset", "2bus_flex", "anm6", "synth30", "synth30s"]


def _resets(g):
    return dict(zip(g["reset_before_step"].tolist(), range(len(g["reset_before_step"]))))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_numpy_port_reproduces_reference_trajectories(seed):
    g, spec = load("anm6easy_traj_seed%d.npz" % seed), anm6easy_spec()
    env, resets = anm_numpy.NumpyANMEnv(spec), _resets(g)
    for t in range(len(g["actions"])):
        if t in resets:
            obs, ok = env.reset(g["reset_s0"][resets[t]])
            assert ok
            np.testing.assert_allclose(obs, g["reset_obs"][resets[t]], rtol=1e-12, atol=1e-12)
        obs, r, term = env.step(g["actions"][t])
        assert term == bool(g["terminated"][t])
        np.testing.assert_allclose(obs, g["obs"][t], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r, g["reward"][t], rtol=1e-12)
        np.testing.assert_allclose(env.state, g["state"][t], rtol=1e-12, atol=1e-12)
        if not term:
            assert env.n_iter == g["n_iter"][t]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c_oracle_reproduces_reference_trajectories(seed):
    g, spec = load("anm6easy_traj_seed%d.npz" % seed), anm6easy_spec()
    env, resets = anm_oracle.OracleEnv(spec, 1), _resets(g)
    sl = spec.full_state_slices()
    for t in range(len(g["actions"])):
        if t in resets:
            k = resets[t]
            obs, state, conv = env.reset(g["reset_s0"][k][None])
            assert conv[0]
            assert rel_err(obs[0], g["reset_obs"][k]) < 1e-9 and rel_err(state[0], g["reset_state"][k]) < 1e-9
        obs, r, term, info = env.step(g["actions"][t][None], want_full=True)
        assert term[0] == g["terminated"][t], (seed, t)
        assert rel_err(obs[0], g["obs"][t]) < 1e-9 and rel_err(r[0], g["reward"][t]) < 1e-9
        assert rel_err(info["state"][0], g["state"][t]) < 1e-9
        assert rel_err(info["e_loss"][0], g["e_loss"][t]) < 1e-8 and rel_err(info["penalty"][0], g["penalty"][t]) < 1e-9
        if not term[0]:
            assert info["n_iter"][0] == g["n_iter"][t], (seed, t)
            for k in ("bus_p", "bus_q", "bus_v_magn", "bus_v_ang", "dev_p", "dev_q", "des_soc", "gen_p_max",
                      "branch_p", "branch_q", "branch_s", "branch_i_magn", "aux"):  # fmt: skip
                assert rel_err(info["full_state"][0][sl[k]], g["full_state"][t][sl[k]]) < 1e-9, (seed, t, k)


@pytest.mark.parametrize("name", NETS)
def test_c_oracle_transitions_and_resets(name):
    g = load("transitions_%s.npz" % name)
    spec = transition_spec(g)
    # host compile: the admittance matrix is bit-identical to Simulator._build_admittance_matrix
    assert np.array_equal(spec.cn.Y_bus_dense.real, g["ybus_re"]) and np.array_equal(spec.cn.Y_bus_dense.imag, g["ybus_im"])
    sl = spec.full_state_slices()
    env = anm_oracle.OracleEnv(spec, 1)
    env.soc[:] = g["soc0"][None]
    for t in range(len(g["reward"])):
        full, r, e, pe, conv, nit = env.transition(g["p_load"][t][None], g["p_pot"][t][None], g["p_set"][t][None], g["q_set"][t][None])
        assert conv[0] == g["converged"][t], (name, t)
        if not conv[0]:
            continue
        assert nit[0] == g["n_iter"][t]
        for k, s in sl.items():
            if k == "aux":
                continue
            a, b = full[0][s], g["full_state"][t][s]
            if k.endswith("_ang"):
                m = np.abs(g["full_state"][t][sl[k.replace("_ang", "_magn")]]) > 1e-6
                a, b = a[m], b[m]
            assert rel_err(a, b, atol=1e-6 if k.startswith("bus_i") else 1e-9) < 1e-8, (name, t, k)
        for x, y in ((r, g["reward"][t]), (e, g["e_loss"][t]), (pe, g["penalty"][t])):
            assert rel_err(x[0], y) < 1e-8
    for k in range(len(g["reset_s0"])):
        env2 = anm_oracle.OracleEnv(spec, 1)
        _, _, conv = env2.reset(g["reset_s0"][k][None])
        assert conv[0] == g["reset_converged"][k]


def test_projection_known_answers():
    """The point->mapped tables of the reference's own tests (test_devices.py:290-291, 556-557)
    and exact box clipping (assertEqual at :541-549), for the Python definition and the C port."""
    from _projection import project_onto_polygon

    tb = load("tables.npz")
    g = load("transitions_2bus_flex.npz")
    spec = transition_spec(g)
    gen, des = spec.cn.devices[1], spec.cn.devices[2]
    Gg = np.array([[-1, 0], [1, 0], [1, 0], [0, -1], [0, 1], [-gen.tau_1, 1], [gen.tau_2, -1]], dtype=float)
    hg = np.array([-gen.p_min, gen.p_max, gen.p_max, -gen.q_min, gen.q_max, gen.rho_1, -gen.rho_2])
    soc, dt = 5.0, 1.0
    Gd = np.array([[-1, 0], [1, 0], [0, -1], [0, 1], [-des.tau_1, 1], [des.tau_2, -1], [des.tau_3, -1], [-des.tau_4, 1],
                   [-1, 0], [1, 0]], dtype=float)  # fmt: skip
    hd = np.array([-des.p_min, des.p_max, -des.q_min, des.q_max, des.rho_1, -des.rho_2, -des.rho_3, des.rho_4,
                   -(soc - des.soc_max) / (dt * des.eff), des.eff * (soc - des.soc_min) / dt])  # fmt: skip
    for pts, mapped, G, h in ((tb["gen_points"], tb["gen_mapped"], Gg, hg), (tb["des_points"], tb["des_mapped"], Gd, hd)):
        rows = np.concatenate([G, h[:, None]], axis=1)
        for p, m in zip(pts / 100, mapped / 100):
            np.testing.assert_allclose(project_onto_polygon(G, h, p), m, atol=1e-12)
            np.testing.assert_allclose(anm_oracle.project(rows, p[0], p[1]), m, atol=1e-12)
    # box clipping is exact, interior points are returned bit-for-bit
    rng = np.random.default_rng(0)
    G = np.array([[-1, 0], [1, 0], [0, -1], [0, 1]], dtype=float)
    h = np.array([0.12, 0.1, 0.3, 0.2])
    rows = np.concatenate([G, h[:, None]], axis=1)
    for p in rng.uniform(-0.5, 0.5, (200, 2)):
        want = np.array([np.clip(p[0], -0.12, 0.1), np.clip(p[1], -0.3, 0.2)])
        assert np.array_equal(project_onto_polygon(G, h, p), want)
        assert np.array_equal(anm_oracle.project(rows, p[0], p[1]), want)


def test_tables_match_reference():
    from gym_anm_b200.networks import anm6easy_tables

    tb = load("tables.npz")
    loads, gens = anm6easy_tables()
    assert np.array_equal(loads, tb["P_loads"]) and np.array_equal(gens, tb["P_maxs"])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present (GPU box)")
def test_numpy_port_bit_exact_vs_live_reference():
    import warnings

    warnings.simplefilter("ignore")
    ref_loader.load_reference()
    from gym_anm.envs import ANM6Easy

    spec = anm6easy_spec()
    ref, mine = ANM6Easy(), anm_numpy.NumpyANMEnv(spec)
    drawn = []
    init = ref.init_state
    ref.init_state = lambda: drawn.append(init()) or drawn[-1]
    obs_ref, _ = ref.reset(seed=123)
    obs, ok = mine.reset(drawn[-1])
    assert ok and np.array_equal(obs, obs_ref)
    rng = np.random.default_rng(5)
    for t in range(150):
        a = rng.uniform(ref.action_space.low, ref.action_space.high)
        o1, r1, t1, _, _ = ref.step(a)
        o2, r2, t2 = mine.step(a)
        assert t1 == t2 and np.array_equal(o1, o2) and r1 == r2, t
        if t1:
            obs_ref, _ = ref.reset()
            obs, ok = mine.reset(drawn[-1])
            assert ok and np.array_equal(obs, obs_ref)


def test_seed_stream_matches_reference_draw_order():
    """init_state draw order integers/uniform/uniform/uniform (+1 date draw after a reset)."""
    spec = anm6easy_spec()
    g = load("anm6easy_traj_seed1.npz")
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(1)))
    for k in range(len(g["reset_s0"])):
        for _ in range(int(g["reset_attempts"][k])):
            s0 = anm_numpy.anm6easy_init_state(spec, rng)
        assert np.array_equal(s0, g["reset_s0"][k])
        rng.integers(1, 365)


@pytest.mark.parametrize("name", ["anm6", "3bus_xfmr"])
def test_list_observation_goldens_c_oracle(name):
    """List-style observation specs with units (anm_env.py:497-549, constants.py:31-48): the host-side spec gives the
    reference's observation-space bounds, and the C restatement replays the recorded episode (same s0 / next_vars /
    actions) to the reference's observations."""
    from golden_util import listobs_noise_mask, listobs_spec

    g = load("listobs_%s.npz" % name)
    spec = listobs_spec(g)
    ok = ~listobs_noise_mask(spec)
    np.testing.assert_allclose(spec.obs_low, g["obs_low"], rtol=1e-12)
    np.testing.assert_allclose(spec.obs_high, g["obs_high"], rtol=1e-12)
    np.testing.assert_allclose(spec.action_low, g["action_low"], rtol=1e-12)
    np.testing.assert_allclose(spec.action_high, g["action_high"], rtol=1e-12)
    env = anm_oracle.OracleEnv(spec, 1)
    resets = {int(t): k for k, t in enumerate(g["reset_before_step"])}
    for t in range(len(g["actions"])):
        if t in resets:
            k = resets[t]
            obs, state, conv = env.reset(g["s0"][k][None])
            assert conv[0]
            assert rel_err(obs[0][ok], g["reset_obs"][k][ok]) < 1e-8, (name, t, "reset obs")
            assert rel_err(state[0], g["reset_state"][k]) < 1e-8
        obs, r, term, info = env.step(g["actions"][t][None], g["next_vars"][t][None])
        assert bool(term[0]) == bool(g["terminated"][t]), (name, t)
        assert rel_err(obs[0][ok], g["obs"][t][ok]) < 1e-8, (name, t, np.abs(obs[0] - g["obs"][t]).argmax())
        assert rel_err(r[0], g["reward"][t]) < 1e-8
        assert rel_err(info["state"][0], g["state"][t]) < 1e-8
