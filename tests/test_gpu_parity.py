"""GPU parity tests: the CUDA path (through the C ABI) against the committed reference
goldens and against the C oracle on identical seeded inputs.

Bars (BASELINE.json north_star): `terminated`, aux and indices bit-exact; voltages, powers,
observations, rewards within 1e-6 relative (we assert 1e-8; typical agreement is 1e-12)."""
import numpy as np
import pytest
import torch

from golden_util import load, rel_err, transition_spec

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _magn_mask(sl, full, key):
    magn_key = key.replace("_ang", "_magn")
    return np.abs(full[..., sl[magn_key]]) > 1e-6


def compare_full(spec, got, want, where=""):
    sl = spec.full_state_slices()
    for k, s in sl.items():
        a, b = got[..., s], want[..., s]
        atol = 1e-9
        if k.endswith("_ang"):
            m = _magn_mask(sl, want, k)
            a, b = a[m], b[m]
        if k in ("bus_i_magn", "bus_i_ang"):
            atol = 1e-6  # injections of zero-injection buses are solver noise (~1e-11)
        err = rel_err(a, b, atol=atol)
        assert err < RTOL * (100 if k in ("bus_i_magn", "bus_i_ang") else 1), (where, k, err)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_anm6easy_golden_trajectory(seed):
    """Replay the reference trajectory (same s0, same actions) through BatchedANM6Easy."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    g = load("anm6easy_traj_seed%d.npz" % seed)
    env = BatchedANM6Easy(1)
    nb = env.native
    resets = dict(zip(g["reset_before_step"].tolist(), range(len(g["reset_before_step"]))))
    full = torch.zeros(1, nb.F, dtype=torch.float64, device=env.device)
    env._extras["full_state"] = full
    for t in range(len(g["actions"])):
        if t in resets:
            k = resets[t]
            obs, state, conv = nb.reset(g["reset_s0"][k][None], obs=env._obs, state=env.state)
            assert bool(conv[0])
            assert rel_err(obs[0].cpu().numpy(), g["reset_obs"][k]) < RTOL
            assert rel_err(state[0].cpu().numpy(), g["reset_state"][k]) < RTOL
            env._term_u8[:] = 0
        obs, r, term, trunc, info = env.step(g["actions"][t][None])
        assert bool(term[0]) == bool(g["terminated"][t]), (seed, t)
        assert rel_err(obs[0].cpu().numpy(), g["obs"][t]) < RTOL, (seed, t)
        assert rel_err(float(r[0]), g["reward"][t]) < RTOL, (seed, t)
        assert rel_err(env.state[0].cpu().numpy(), g["state"][t]) < RTOL, (seed, t)
        assert rel_err(float(env.e_loss[0]), g["e_loss"][t]) < 1e-7, (seed, t)
        assert rel_err(float(env.penalty[0]), g["penalty"][t]) < RTOL, (seed, t)
        if not g["terminated"][t]:
            assert int(env.n_iter[0]) == int(g["n_iter"][t]), (seed, t)
            compare_full(env.spec, full[0].cpu().numpy(), g["full_state"][t], where=(seed, t))


def test_anm6easy_seeded_reset_matches_reference():
    """reset(seed=0) over 3 envs: env i uses PCG64(SeedSequence(i)) like the reference run with seed=i."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    env = BatchedANM6Easy(3)
    obs, _ = env.reset(seed=0)
    for i in range(3):
        g = load("anm6easy_traj_seed%d.npz" % i)
        assert rel_err(obs[i].cpu().numpy(), g["reset_obs"][0]) < RTOL
    # first 20 steps of every env follow its golden trajectory
    for t in range(20):
        a = np.stack([load("anm6easy_traj_seed%d.npz" % i)["actions"][t] for i in range(3)])
        obs, r, term, _, _ = env.step(a)
        for i in range(3):
            g = load("anm6easy_traj_seed%d.npz" % i)
            assert rel_err(obs[i].cpu().numpy(), g["obs"][t]) < RTOL
            assert rel_err(float(r[i]), g["reward"][t]) < RTOL


def test_single_env_facade_reference_shapes():
    from gym_anm_b200.anm6 import ANM6Easy

    g = load("anm6easy_traj_seed0.npz")
    env = ANM6Easy()
    obs, info = env.reset(seed=0)
    assert isinstance(obs, np.ndarray) and obs.shape == (18,) and info == {}
    assert rel_err(obs, g["reset_obs"][0]) < RTOL
    o, r, term, trunc, info = env.step(g["actions"][0])
    assert isinstance(r, float) and isinstance(term, bool) and trunc is False and info == {}
    assert rel_err(o, g["obs"][0]) < RTOL and rel_err(r, g["reward"][0]) < RTOL
    with pytest.raises(AssertionError):
        env.step(np.array([1e3, 0, 0, 0, 0, 0.0]))


@pytest.mark.parametrize("name", ["2bus", "3bus_loop", "3bus_xfmr", "3bus_reset", "2bus_flex", "anm6", "synth30", "synth30s"])
def test_transition_goldens(name):
    """Simulator.transition sequences recorded from the reference (8/16/32-lane kernels)."""
    from gym_anm_b200.native import NativeBatch

    g = load("transitions_%s.npz" % name)
    spec = transition_spec(g)
    nb = NativeBatch(spec, 1)
    nb.set_state(soc=g["soc0"][None])
    for t in range(len(g["reward"])):
        full, r, e, pe, conv = nb.transition(g["p_load"][t][None], g["p_pot"][t][None], g["p_set"][t][None], g["q_set"][t][None])
        assert bool(conv[0]) == bool(g["converged"][t]), (name, t)
        if not g["converged"][t]:
            continue
        compare_full(spec, np.concatenate([full[0].cpu().numpy(), np.zeros(spec.K)]),
                     np.concatenate([g["full_state"][t], np.zeros(spec.K)]), where=(name, t))  # fmt: skip
        for x, y in ((r, g["reward"][t]), (e, g["e_loss"][t]), (pe, g["penalty"][t])):
            assert rel_err(float(x[0]), y, atol=1e-9) < 1e-7, (name, t)
    if name == "2bus_flex":  # known-answer tables of the reference's own device tests
        tb = load("tables.npz")
        sl = spec.full_state_slices()
        nb2 = NativeBatch(spec, 1)
        nb2.set_state(soc=np.array([[5.0]]))
        for k in range(len(tb["gen_points"])):
            gp, dp = tb["gen_points"][k], tb["des_points"][k % len(tb["des_points"])]
            nb2.set_state(soc=np.array([[5.0]]))
            full, *_ = nb2.transition(np.zeros((1, 0)), np.array([[10.0]]), np.array([[gp[0], dp[0]]]), np.array([[gp[1], dp[1]]]))
            row = full[0].cpu().numpy()
            np.testing.assert_allclose(row[sl["dev_p"]][1] * 100, tb["gen_mapped"][k][0], rtol=0, atol=1e-9)
            np.testing.assert_allclose(row[sl["dev_q"]][1] * 100, tb["gen_mapped"][k][1], rtol=0, atol=1e-9)
            dm = tb["des_mapped"][k % len(tb["des_points"])]
            np.testing.assert_allclose(row[sl["dev_p"]][2] * 100, dm[0], rtol=0, atol=1e-9)
            np.testing.assert_allclose(row[sl["dev_q"]][2] * 100, dm[1], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["3bus_reset", "anm6", "synth30", "synth30s", "2bus", "3bus_loop", "3bus_xfmr", "2bus_flex"])
def test_reset_goldens(name):
    """Simulator.reset cases recorded from the reference (simulator.py:225-293): the converged flags, and for the
    converged cases the whole electrical state after the reset plus the state / observation vectors derived from it."""
    from gym_anm_b200.native import NativeBatch

    g = load("transitions_%s.npz" % name)
    spec = transition_spec(g)
    n = len(g["reset_s0"])
    nb = NativeBatch(spec, n)
    full = nb.empty(n, nb.F)
    nb.set_reset_full_state(full)
    obs, state, conv = nb.reset(g["reset_s0"])
    conv = conv.cpu().numpy().astype(bool)
    assert np.array_equal(conv, g["reset_converged"])
    assert conv.any()
    full, obs, state = full.cpu().numpy(), obs.cpu().numpy(), state.cpu().numpy()
    sl, cn, m = spec.full_state_slices(), spec.cn, spec.cn.baseMVA
    pos = {d: k for k, d in enumerate(cn.devices)}
    for k in np.flatnonzero(conv):
        want = g["reset_full_state"][k]
        compare_full(spec, full[k], np.concatenate([want, np.zeros(spec.K)]), where=(name, k))
        # the state vector (anm_env.py:562-592): dev_p MW, dev_q MVAr, des_soc MWh, gen_p_max MW
        vec = np.concatenate([want[sl["dev_p"]] * m, want[sl["dev_q"]] * m,
                              [want[sl["des_soc"]][pos[i]] * m for i in cn.des_ids],
                              [want[sl["gen_p_max"]][pos[i]] * m for i in cn.gen_ids]])  # fmt: skip
        assert rel_err(state[k], vec) < RTOL, (name, k)
        assert rel_err(obs[k], np.clip(vec, spec.obs_low, spec.obs_high)) < RTOL, (name, k)


@pytest.mark.parametrize("name", ["anm6", "3bus_xfmr"])
def test_list_observation_goldens(name):
    """List-style observation specs in every unit (kV / kA / degree / MVA / MWh / rad / p.u.; anm_env.py:497-549,
    constants.py:31-48; the pattern of the reference's tests/envs/custom_obs_space.py:33-72) through the kernel's
    observation map (unit divisors, angle gathers, magnitudes): the reference's reset and step observations for the
    recorded s0 / next_vars / actions.  Four identical instances: every lane group must give the same row."""
    from golden_util import listobs_noise_mask, listobs_spec
    from gym_anm_b200.native import NativeBatch

    g = load("listobs_%s.npz" % name)
    spec = listobs_spec(g)
    ok = ~listobs_noise_mask(spec)
    B = 4
    nb = NativeBatch(spec, B)
    rep = lambda a: np.repeat(np.asarray(a)[None], B, axis=0)  # noqa: E731
    resets = {int(t): k for k, t in enumerate(g["reset_before_step"])}
    state = nb.empty(B, nb.S)
    for t in range(len(g["actions"])):
        if t in resets:
            k = resets[t]
            obs, st, conv = nb.reset(rep(g["s0"][k]))
            assert bool(conv.all())
            assert rel_err(obs.cpu().numpy()[:, ok], rep(g["reset_obs"][k])[:, ok]) < RTOL, (name, t, "reset")
            assert rel_err(st.cpu().numpy(), rep(g["reset_state"][k])) < RTOL
        obs, r, term = nb.step(rep(g["actions"][t]), rep(g["next_vars"][t]), extras={"state": state})
        assert np.array_equal(term.cpu().numpy().astype(bool), rep(bool(g["terminated"][t]))), (name, t)
        o = obs.cpu().numpy()
        assert np.array_equal(o, rep(o[0]))
        bad = np.abs(o[0] - g["obs"][t])
        assert rel_err(o[:, ok], rep(g["obs"][t])[:, ok]) < RTOL, (name, t, int(np.argmax(bad * ok)))
        assert rel_err(r.cpu().numpy(), rep(g["reward"][t])) < RTOL
        assert rel_err(state.cpu().numpy(), rep(g["state"][t])) < RTOL


@pytest.mark.parametrize("B,steps", [(4096, 120)])
def test_batch_vs_oracle(B, steps):
    """BASELINE config 2 (4096 envs, fp64): CUDA vs the C oracle on identical seeded inputs."""
    import anm_oracle
    from gym_anm_b200.anm6 import BatchedANM6Easy

    env = BatchedANM6Easy(B, validate_actions=False)
    spec = env.spec
    rng = np.random.default_rng(2020)
    # s0: valid initial states from the env's own init_state with per-env PCG64 streams
    env._seed_rngs(2020)
    s0 = env.init_state_batch(np.arange(B))
    cpu = anm_oracle.OracleEnv(spec, B)
    obs_c, state_c, conv_c = cpu.reset(s0)
    obs_g, state_g, conv_g = env.native.reset(s0, obs=env._obs, state=env.state)
    assert np.array_equal(conv_g.cpu().numpy().astype(bool), conv_c)
    assert rel_err(obs_g.cpu().numpy(), obs_c) < RTOL
    env._term_u8.copy_(torch.as_tensor((~conv_c).astype(np.uint8)))
    n_term = 0
    for t in range(steps):
        a = rng.uniform(spec.action_low, spec.action_high, size=(B, 6))
        obs_g, r_g, term_g, _, _ = env.step(a)
        obs_c, r_c, term_c, info = cpu.step(a)
        tg = term_g.cpu().numpy()
        assert np.array_equal(tg, term_c), (t, int((tg != term_c).sum()))
        assert rel_err(obs_g.cpu().numpy(), obs_c) < RTOL, t
        assert rel_err(r_g.cpu().numpy(), r_c) < RTOL, t
        ok = ~term_c
        assert np.array_equal(env.n_iter.cpu().numpy()[ok], info["n_iter"][ok]), t
        n_term += int(term_c.sum())
        if t % 10 == 9 and t + 1 < steps:  # bring the terminated instances back (same s0 on both sides)
            m = term_c.astype(np.uint8)
            cpu.reset(s0, mask=m)
            env.native.reset(s0, mask=m, obs=env._obs, state=env.state)
            env._term_u8.copy_(torch.as_tensor(cpu.terminated))
    soc_g, aux_g, _ = env.native.get_state()
    assert np.array_equal(aux_g.cpu().numpy(), cpu.aux)          # aux: bit-exact
    assert rel_err(soc_g.cpu().numpy()[~term_c], cpu.soc[~term_c]) < RTOL
    assert n_term > 0  # the divergent cases were exercised


def _synth30_setup(B, impedance_scale, seed):
    import anm_oracle
    from gym_anm_b200.env_spec import HostEnvSpec
    from gym_anm_b200.native import NativeBatch
    from gym_anm_b200.networks import synth_feeder_network

    spec = HostEnvSpec(synth_feeder_network(impedance_scale=impedance_scale), "state", 1, 0.25, 0.99, 100,
                       np.array([[0, 95]]), (1, 100))  # fmt: skip
    cn = spec.cn
    rng = np.random.default_rng(seed)
    nb, cpu = NativeBatch(spec, B), anm_oracle.OracleEnv(spec, B)
    D, ns = cn.N_device, cn.N_des
    s0 = np.zeros((B, spec.state_N))
    pos = {d: k for k, d in enumerate(cn.devices)}
    f = 1.0 if impedance_scale == 1.0 else 0.3  # a weak grid needs a light initial state to start from
    for i in cn.load_ids:
        s0[:, pos[i]] = rng.uniform(cn.devices[i].p_min * 100 * f, 0, B)
    for k, i in enumerate(cn.gen_ids):
        s0[:, pos[i]] = rng.uniform(0, cn.devices[i].p_max * 100 * f, B)
        s0[:, 2 * D + ns + k] = rng.uniform(cn.devices[i].p_max * 100 * f, cn.devices[i].p_max * 100, B)
    for k, i in enumerate(cn.des_ids):
        s0[:, 2 * D + k] = rng.uniform(0, cn.devices[i].soc_max * 100, B)
    return spec, cn, rng, nb, cpu, s0


def _synth30_inputs(spec, cn, rng, B, t):
    a = rng.uniform(spec.action_low, spec.action_high, size=(B, len(spec.action_low)))
    nv = np.concatenate([rng.uniform([cn.devices[i].p_min * 100 for i in cn.load_ids], 0, (B, cn.N_load)),
                         rng.uniform(0, [cn.devices[i].p_max * 100 for i in cn.gen_ids], (B, cn.N_non_slack_gen)),
                         np.full((B, 1), float(t))], axis=1)  # fmt: skip
    return a, nv


def test_synth30_env_vs_oracle():
    """BASELINE config 4 (30-bus, 32-lane kernel, caller-supplied next_vars)."""
    B = 256
    spec, cn, rng, nb, cpu, s0 = _synth30_setup(B, 1.0, 30)
    obs_c, state_c, conv_c = cpu.reset(s0)
    obs_g, state_g, conv_g = nb.reset(s0)
    assert np.array_equal(conv_g.cpu().numpy().astype(bool), conv_c) and conv_c.all()
    assert rel_err(obs_g.cpu().numpy(), obs_c) < RTOL
    for t in range(10):
        a, nv = _synth30_inputs(spec, cn, rng, B, t)
        obs_g, r_g, term_g = nb.step(a, nv)
        obs_c, r_c, term_c, info = cpu.step(a, nv)
        assert np.array_equal(term_g.cpu().numpy().astype(bool), term_c)
        assert rel_err(obs_g.cpu().numpy(), obs_c) < RTOL and rel_err(r_g.cpu().numpy(), r_c) < RTOL


def test_synth30_stressed_full_size_vs_oracle():
    """BASELINE config 4 at its full size (8192 instances) on the stressed 30-bus feeder (four times the branch
    impedances; the reference's own numbers for it are in transitions_synth30s.npz): a few per cent of the solves
    diverge (100 iterations through the block-sparse solver), others need 5-7 iterations.  `terminated` and the
    Newton iteration counts of every converged solve are bit-exact against the pivoting C oracle."""
    B = 8192
    spec, cn, rng, nb, cpu, s0 = _synth30_setup(B, 4.0, 31)
    obs_c, state_c, conv_c = cpu.reset(s0)
    obs_g, state_g, conv_g = nb.reset(s0)
    assert np.array_equal(conv_g.cpu().numpy().astype(bool), conv_c) and conv_c.mean() > 0.99
    n_term, hist = 0, np.zeros(102, dtype=np.int64)
    nit = nb.empty(B, dtype=torch.int32)
    for t in range(3):
        a, nv = _synth30_inputs(spec, cn, rng, B, t)
        obs_g, r_g, term_g = nb.step(a, nv, extras={"n_iter": nit})
        obs_c, r_c, term_c, info = cpu.step(a, nv)
        tg = term_g.cpu().numpy().astype(bool)
        assert np.array_equal(tg, term_c), (t, int((tg != term_c).sum()))
        live = ~term_c
        assert rel_err(obs_g.cpu().numpy()[live], obs_c[live]) < RTOL and rel_err(r_g.cpu().numpy(), r_c) < RTOL
        ran = live & (info["n_iter"] > 0)  # instances that were already terminal report 0 iterations
        assert np.array_equal(nit.cpu().numpy()[ran], info["n_iter"][ran])
        hist += np.bincount(np.minimum(info["n_iter"][ran], 101), minlength=102)
        n_term += int(term_c.sum())
        if t == 0:
            assert term_c.mean() > 0.01, "the stress fixture must contain divergent instances"
    assert hist[5:9].sum() > 0, "the stress fixture must contain 5-8 iteration solves"


def _singular_leaf_network(bsh):
    """0 - 1 - 2 chain whose leaf line (r = x = 0.5 p.u.) carries the line charging `bsh`: with bsh = 2.0 the leaf's
    2x2 Jacobian block at the flat start is [[1, 1], [-1, -1]] -- exactly singular -- while the whole Jacobian is
    regular and well conditioned (cond ~ 54)."""
    N_ = None
    slack = [0, 0, 0, N_, 200, -200, 200, -200, N_, N_, N_, N_, N_, N_, N_]
    return {"baseMVA": 1,
            "bus": np.array([[0, 0, 50, 1.0, 1.0], [1, 1, 50, 1.5, 0.5], [2, 1, 50, 1.5, 0.5]]),
            "branch": np.array([[0, 1, 0.01, 0.1, 0.0, 30, 1, 0], [1, 2, 0.5, 0.5, bsh, 30, 1, 0]]),
            "device": np.array([slack, [1, 1, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_],
                                [2, 2, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_]], dtype=object)}  # fmt: skip


_SINGULAR_SCRIPT = r"""
import os, sys, json
import numpy as np
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "oracle"), os.path.join(%(root)r, "tests")]
os.environ["ANM_DEBUG_NR_MAXIT"] = "%(maxit)d"
if "%(solver)s":
    os.environ["ANM_SOLVER"] = "%(solver)s"
import anm_oracle
from gym_anm_b200.env_spec import HostEnvSpec
from gym_anm_b200.native import NativeBatch
from test_gpu_parity import _singular_leaf_network
out = {}
for bsh in (2.0, 2.0 * (1 + 1e-13), 2.0 * (1 - 1e-10), 1.0):
    spec = HostEnvSpec(_singular_leaf_network(bsh), "state", 0, 0.5, 0.9, 100)
    pl = np.array([[-0.3, -0.01], [-0.1, -0.02], [0.0, 0.0], [-1.0, -0.005], [-0.2, -0.3]])
    B = len(pl)
    z = np.zeros((B, 0))
    nb, cpu = NativeBatch(spec, B), anm_oracle.OracleEnv(spec, B)
    full_g, r_g, e_g, pe_g, conv_g = nb.transition(pl, z, z, z)
    full_c, r_c, e_c, pe_c, conv_c, nit_c = cpu.transition(pl, z, z, z)
    sl = spec.full_state_slices()
    vg, vc = full_g.cpu().numpy()[:, sl["bus_v_magn"]], full_c[:, sl["bus_v_magn"]]
    ag, ac = full_g.cpu().numpy()[:, sl["bus_v_ang"]], full_c[:, sl["bus_v_ang"]]
    out[repr(bsh)] = dict(conv_g=conv_g.cpu().numpy().astype(bool).tolist(), conv_c=conv_c.tolist(),
                          v_err=float(np.nanmax(np.abs(vg - vc) / np.maximum(np.abs(vc), 1e-300))),
                          a_err=float(np.nanmax(np.abs(ag - ac))), finite=bool(np.isfinite(vg).all() and np.isfinite(vc).all()),
                          vmax=float(np.abs(vc).max()))
print("RESULT" + json.dumps(out))
"""


@pytest.mark.parametrize("solver", ["", "sparse"])
@pytest.mark.parametrize("maxit", [1, 2, 100])
def test_singular_schur_block_guard_matches_pivoting_oracle(maxit, solver):
    """Adversarial case for the tree elimination (RadialNR inverts 2x2 blocks without pivoting across blocks; the
    reference's SuperLU pivots): the leaf block of the first Newton iteration is exactly / nearly singular while the
    Jacobian is regular.  The singular-block guard must redo that iteration with the dense partially pivoted solver:
    after ONE iteration (ANM_DEBUG_NR_MAXIT=1, in a subprocess) the iterate is finite and equals the pivoting C oracle's
    (an unguarded block elimination gives inf / NaN here); after two iterations still; and the full solve takes the
    same decision.  bsh = 1.0 is the regular control case.  `solver`: the default (radial tree elimination) and the
    block-sparse LU (ANM_SOLVER=sparse), whose guard redoes the solve on a dense system in global scratch memory."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _SINGULAR_SCRIPT % {"root": root, "maxit": maxit, "solver": solver}], capture_output=True,
                         text=True, timeout=600)  # fmt: skip
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    for bsh, o in out.items():
        assert o["conv_g"] == o["conv_c"], (maxit, solver, bsh, o)
        if maxit <= 2:
            assert o["finite"], (maxit, bsh, o)
            assert o["v_err"] < 1e-8 and o["a_err"] < 1e-8, (maxit, bsh, o)
        elif float(bsh) == 1.0:
            assert all(o["conv_c"]) and o["v_err"] < 1e-8


def test_tensor_level_hooks_of_a_custom_env():
    """A custom environment in the style of the reference's examples/simple_env.py:33 (2-bus grid, random load and
    auxiliary variable per step, callable observation) written once with the reference's per-environment hooks
    (`next_vars(s_t)`, a callable `observation`) and once with the tensor-level hooks (`next_vars_batch`,
    `observation_batch`: torch ops on the [B, .] CUDA tensors, no host round trip): identical trajectories."""
    from gym_anm_b200.anm_env import BatchedANMEnv

    N_ = None
    network = {
        "baseMVA": 100,
        "bus": np.array([[0, 0, 132, 1.0, 1.0], [1, 1, 33, 1.1, 0.9]]),
        "device": np.array([[0, 0, 0, N_, 200, -200, 200, -200, N_, N_, N_, N_, N_, N_, N_],
                            [1, 1, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_],
                            [2, 1, 2, N_, 30, 0, 30, -30, N_, N_, N_, N_, N_, N_, N_]], dtype=object),
        "branch": np.array([[0, 1, 0.01, 0.1, 0.0, 3, 1, 0]]),
    }
    B, T = 512, 12
    draws = np.random.default_rng(3).uniform(size=(T + 1, B, 2))  # the "random" hooks replay these

    def obs_fn(s):  # a callable observation: scaled load power and the auxiliary variable
        return np.array([s[1] * 0.125, s[-1]])

    class PerEnv(BatchedANMEnv):
        def __init__(self, observation):
            super().__init__(network, observation, 1, 0.25, 0.9, 100, np.array([[0, 10]]), (1, 100), 1, num_envs=B)
            self.k = 0

        def init_state(self):
            u = draws[0, self._env_index]
            s0 = np.zeros(self.state_N)  # [dev_p x3, dev_q x3, gen_p_max, aux]
            s0[1], s0[4] = -5 * u[0], -u[0]
            s0[2], s0[5], s0[6] = 10 * u[1], 2 * u[1] - 1, 20 * u[1]
            return s0

        def next_vars(self, s_t):
            i = self._env_index
            return np.array([-10 * draws[self.k, i, 0], 30 * draws[self.k, i, 1], np.floor(10 * draws[self.k, i, 1])])

    class Tensor(PerEnv):
        def next_vars_batch(self, state):
            d = torch.as_tensor(draws[self.k], device=self.device)
            return torch.stack([-10 * d[:, 0], 30 * d[:, 1], torch.floor(10 * d[:, 1])], dim=1)

        def observation_batch(self, state):
            return torch.stack([state[:, 1] * 0.125, state[:, -1]], dim=1)

    envs = [PerEnv(obs_fn), Tensor(obs_fn)]
    outs = []
    for env in envs:
        obs, _ = env.reset()
        rows = [obs.clone()]
        rng = np.random.default_rng(5)
        for t in range(T):
            env.k = t + 1
            a = rng.uniform(env.action_space.low, env.action_space.high, size=(B, env.action_space.shape[0]))
            o, r, d, _, _ = env.step(a)
            rows += [o.clone(), r.clone(), d.double()]
        outs.append(rows)
    assert envs[0].observation_space.shape == (2,)
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    assert float(outs[1][-3].abs().sum()) > 0


def test_fast_math_ulp_sweep():
    """The kernel's own math routines against libm (NumPy): `sincos_fast` (V = |V| e^{j theta} in the Newton loop),
    the sqrt(fma) magnitude (|V|, |S|, |I|) and the RCP64H + two Newton steps reciprocal (2x2 block inverses)."""
    import ctypes as C

    from gym_anm_b200 import _capi

    lib = _capi.load_library()
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = np.random.default_rng(7)

    def run(kind, x, y=None):
        xd = torch.as_tensor(x, device=dev)
        yd = None if y is None else torch.as_tensor(y, device=dev)
        a, b = torch.empty_like(xd), torch.empty_like(xd)
        _capi.check(lib.anm_debug_math(kind, xd.numel(), C.c_void_p(xd.data_ptr()),
                                       None if yd is None else C.c_void_p(yd.data_ptr()), C.c_void_p(a.data_ptr()),
                                       C.c_void_p(b.data_ptr()), None), lib)  # fmt: skip
        torch.cuda.synchronize()
        return a.cpu().numpy(), b.cpu().numpy()

    def ulps(got, want):
        return np.abs(got - want) / np.spacing(np.abs(want))

    n = 400000
    # sincos: angles of converging iterations (|theta| < 4), of wandering ones (up to 1e5: libdevice's fast-path range)
    for scale, max_ulp in ((4.0, 2.0), (100.0, 2.0), (1e5, 2.0)):
        x = rng.uniform(-scale, scale, n)
        sn, cs = run(0, x)
        big = np.abs(np.sin(x)) > 1e-3  # relative accuracy away from the zeros; absolute accuracy everywhere
        assert ulps(sn[big], np.sin(x)[big]).max() <= max_ulp
        big = np.abs(np.cos(x)) > 1e-3
        assert ulps(cs[big], np.cos(x)[big]).max() <= max_ulp
        assert np.abs(sn - np.sin(x)).max() < 3e-16 and np.abs(cs - np.cos(x)).max() < 3e-16
    # divergent trajectories: |theta| up to 2^50 stays on the branch-free path with an absolute error of a few 1e-16
    x = rng.uniform(-1.0, 1.0, n) * 10.0 ** rng.uniform(5, 15, n)
    x = x[np.abs(x) < 2.0**50]
    sn, cs = run(0, x)
    assert np.abs(sn - np.sin(x)).max() < 1e-15 and np.abs(cs - np.cos(x)).max() < 1e-15
    # beyond 2^50, inf, NaN: libdevice
    x = np.array([2.0**51, -2.0**60, 1e300, np.inf, -np.inf, np.nan, 0.0, -0.0])
    sn, cs = run(0, x)
    with np.errstate(invalid="ignore"):
        np.testing.assert_allclose(sn, np.sin(x), rtol=1e-15, atol=1e-15, equal_nan=True)
        np.testing.assert_allclose(cs, np.cos(x), rtol=1e-15, atol=1e-15, equal_nan=True)
    # magnitudes: per-unit ranges and far beyond; hypot is the reference (numpy abs of a complex)
    x = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-8, 8, n)
    y = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-8, 8, n)
    mag, _ = run(1, x, y)
    assert ulps(mag, np.hypot(x, y)).max() <= 1.0
    # reciprocal
    x = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-30, 30, n)
    x = x[x != 0.0]
    rc, _ = run(2, x)
    assert ulps(rc, 1.0 / x).max() <= 1.0


def test_host_buffer_path_and_terminal_semantics():
    """anm_step_host (H2D + step + D2H in the library) == device path; terminated envs stay
    terminated with zeros / 0.0 / True until reset (anm_env.py:365-367)."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 512
    a_env, b_env = BatchedANM6Easy(B, validate_actions=False), BatchedANM6Easy(B, validate_actions=False)
    a_env.reset(seed=7)
    b_env.reset(seed=7)
    rng = np.random.default_rng(1)
    obs_h = torch.zeros(B, 18, dtype=torch.float64).pin_memory()
    r_h = torch.zeros(B, dtype=torch.float64).pin_memory()
    t_h = torch.zeros(B, dtype=torch.uint8).pin_memory()
    was_term = np.zeros(B, dtype=bool)
    for t in range(60):
        a = rng.uniform(a_env.spec.action_low, a_env.spec.action_high, size=(B, 6))
        obs, r, term, _, _ = a_env.step(a)
        a_pin = torch.as_tensor(a).pin_memory()
        b_env.native.step_host(a_pin, None, obs_h, r_h, t_h)
        assert torch.equal(obs.cpu(), obs_h) and torch.equal(r.cpu(), r_h) and torch.equal(term.cpu(), t_h.bool())
        tn = term.cpu().numpy()
        assert np.all(tn[was_term]) and np.all(r.cpu().numpy()[was_term] == 0.0)
        assert np.all(obs.cpu().numpy()[was_term] == 0.0)
        newly = tn & ~was_term
        assert np.all(r.cpu().numpy()[newly] == -100 / (1 - 0.995))
        was_term = tn
    assert was_term.any()


def test_autoreset_pool():
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 1024
    env = BatchedANM6Easy(B, validate_actions=False)
    env.reset(seed=11)
    pool = env.state.clone()  # reconstructed states of converged resets are valid s0 rows
    env.native.set_autoreset_pool(pool)
    rng = np.random.default_rng(3)
    prev_term = np.zeros(B, dtype=bool)
    resets = 0
    for t in range(80):
        a = rng.uniform(env.spec.action_low, env.spec.action_high, size=(B, 6))
        obs, r, term, _, _ = env.step(a)
        tn, rn = term.cpu().numpy(), r.cpu().numpy()
        assert not np.any(tn & prev_term)           # a terminated env is re-initialised on the next step
        assert np.all(rn[prev_term] == 0.0)          # ... and that step returns reward 0
        assert np.all(np.abs(obs.cpu().numpy()[prev_term]).sum(1) > 0)
        resets += int(prev_term.sum())
        prev_term = tn
    assert resets > 0


def test_state_dict_roundtrip_and_simulator_view():
    from gym_anm_b200.anm6 import BatchedANM6Easy

    env = BatchedANM6Easy(64, validate_actions=False)
    env.reset(seed=5)
    rng = np.random.default_rng(9)
    acts = rng.uniform(env.spec.action_low, env.spec.action_high, size=(6, 64, 6))
    for t in range(3):
        env.step(acts[t])
    ck = env.state_dict()
    ref = [tuple(x.clone() for x in env.step(acts[t])[:3]) for t in range(3, 6)]
    env.load_state_dict(ck)
    for t in range(3, 6):
        o, r, term, _, _ = env.step(acts[t])
        assert torch.equal(o, ref[t - 3][0]) and torch.equal(r, ref[t - 3][1]) and torch.equal(term, ref[t - 3][2])
    sim = env.simulator
    assert sim.N_bus == 6 and sim.N_device == 7 and sim.Y_bus.shape == (6, 6)
    full, r, e, pe, conv = sim.transition({1: -5.0, 3: -10.0, 5: -20.0}, {2: 20.0, 4: 30.0}, {2: 10.0, 4: 20.0, 6: 5.0},
                                          {2: 1.0, 4: 2.0, 6: 0.0})  # fmt: skip
    d = sim.state_dict(0)
    assert set(d) >= {"bus_v_magn", "dev_p", "branch_s", "des_soc", "gen_p_max"} and abs(d["bus_v_magn"]["pu"][0] - 1.0) < 1e-12


def test_single_env_simulator_state_and_render_bridge():
    """`ANM6Easy().simulator.state` -- the nested dict the reference's renderer (anm6.py:101-109) and its MPC agents
    (mpc.py:419, mpc_constant.py:23-29) read -- against the reference's recorded full state, step by step; and the
    render bridge: `render()` feeds a renderer object the reference's start / update call sequence and payload."""
    from gym_anm_b200.anm6 import ANM6Easy

    g = load("anm6easy_traj_seed0.npz")

    class Recorder:  # the three functions of gym_anm/envs/anm6_env/rendering/py/rendering.py
        def __init__(self):
            self.calls = []

        def start(self, title, dev_type, ps, qs, branch_rate, bus_v_min, bus_v_max, soc_max, costs_range):
            self.calls.append(("start", title, list(dev_type), list(ps), list(qs), list(branch_rate), list(soc_max), costs_range))
            return "http", type("Ws", (), {"address": "ws://test"})()

        def update(self, address, date, year_count, dev_p, dev_q, branch_s, des_soc, gen_p_max, bus_v_magn, costs, collapsed):
            self.calls.append(("update", address, date, year_count, list(dev_p), list(dev_q), list(branch_s), list(des_soc),
                               list(gen_p_max), list(bus_v_magn), list(costs), collapsed))

        def close(self, http, ws):
            self.calls.append(("close",))

    rec = Recorder()
    env = ANM6Easy(renderer=rec)
    obs, _ = env.reset(seed=0)
    assert rel_err(obs, g["reset_obs"][0]) < RTOL
    st = env.simulator.state
    assert set(st) == {"bus_p", "bus_q", "bus_v_magn", "bus_v_ang", "bus_i_magn", "bus_i_ang", "dev_p", "dev_q", "des_soc",
                       "gen_p_max", "branch_p", "branch_q", "branch_s", "branch_i_magn", "branch_i_ang"}
    assert list(st["des_soc"]["MWh"]) == [6] and list(st["gen_p_max"]["MW"]) == [2, 4]
    assert list(st["branch_s"]["MVA"]) == [(0, 1), (1, 2), (1, 3), (2, 4), (2, 5)]
    np.testing.assert_allclose([st["dev_p"]["MW"][i] for i in range(7)], g["reset_state"][0][:7], rtol=1e-9, atol=1e-9)
    env.render()
    assert rec.calls[0][0] == "start" and rec.calls[0][1] == "ANM6Easy" and rec.calls[0][2] == [0, -1, 2, -1, 2, -1, 3]
    assert rec.calls[0][5] == [32.0, 25.0, 18.0, 18.0, 18.0] and rec.calls[0][7] == (1, 100)
    from gym_anm_b200.env_spec import anm6easy_spec

    sl = anm6easy_spec().full_state_slices()
    n_upd = 1
    for t in range(40):
        if g["terminated"][t]:
            break
        obs, r, term, trunc, info = env.step(g["actions"][t])
        assert rel_err(obs, g["obs"][t]) < RTOL and abs(r - g["reward"][t]) <= 1e-8 * max(1.0, abs(g["reward"][t]))
        st, want = env.simulator.state, g["full_state"][t]
        np.testing.assert_allclose(list(st["bus_v_magn"]["pu"].values()), want[sl["bus_v_magn"]], rtol=1e-9)
        np.testing.assert_allclose(list(st["branch_s"]["MVA"].values()), want[sl["branch_s"]] * 100, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(list(st["dev_q"]["MVAr"].values()), want[sl["dev_q"]] * 100, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st["des_soc"]["pu"][6], want[sl["des_soc"]][6], rtol=1e-9, atol=1e-12)
        env.render(skip_frames=1)  # every second call updates
        if (t + 1) % 2 == 0:
            n_upd += 1
            u = rec.calls[-1]
            assert u[0] == "update" and u[1] == "ws://test" and u[2] == env.date and not u[-1]
            np.testing.assert_allclose(u[4], want[sl["dev_p"]] * 100, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(u[9], want[sl["bus_v_magn"]], rtol=1e-9)
            assert u[10] == [env.e_loss, env.penalty]
    assert sum(c[0] == "update" for c in rec.calls) == n_upd
    env.close()
    assert rec.calls[-1] == ("close",)


def test_checkpoint_restores_costs_observation_and_random_streams():
    """state_dict / load_state_dict: e_loss / penalty / observation of the last step and the random
    streams -- host Generators and the device-side PCG64 streams of `device_init=True` -- so that the resets AFTER a
    resume draw the same initial states."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    for device_init in (False, True):
        env = BatchedANM6Easy(128, validate_actions=False, device_init=device_init)
        env.reset(seed=11)
        rng = np.random.default_rng(2)
        acts = rng.uniform(env.spec.action_low, env.spec.action_high, size=(6, 128, 6))
        for t in range(3):
            env.step(acts[t])
        ck = env.state_dict()
        saved = {k: env.__dict__[k].clone() for k in ("e_loss", "penalty", "_obs")}

        def continue_run(e):
            out = []
            for t in range(3, 6):
                o, r, d, _, _ = e.step(acts[t])
                out += [o.clone(), r.clone(), d.clone()]
                if bool(d.any()):
                    o2, _ = e.reset(mask=d)  # draws from the streams: must continue where the checkpoint left them
                    out.append(o2.clone())
            o3, _ = e.reset()  # and a full reset without a new seed
            return out + [o3.clone()]

        want = continue_run(env)
        env2 = BatchedANM6Easy(128, validate_actions=False, device_init=device_init)
        env2.reset(seed=999)  # a different history, then the checkpoint
        env2.load_state_dict(ck)
        for k, v in saved.items():
            assert torch.equal(env2.__dict__[k], v), k
        got = continue_run(env2)
        assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want)), device_init


def test_launch_ordinals_survive_the_32_bit_wrap():
    """Launch chaining derives the ordinal of a launch from a device-side ticket counter (ticket / grid + 1).  With a
    grid that is not a power of two a 32-bit counter would split a launch in two ordinals when it wraps (~2^32 CTA
    starts: minutes of closed-loop stepping) and dead-lock the next chained wait; the counter is 64 bits wide and the
    ordinals are compared modulo 2^32.  Presets the handle just below both wraps (the ticket's 2^32 and the ordinal's
    2^32) and runs chained steps, a rollout and a CUDA-graph replay across them against an untouched twin."""
    import ctypes as C

    from gym_anm_b200 import _capi
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 1000  # 250 CTAs: not a power of two
    rng = np.random.default_rng(4)
    acts = torch.as_tensor(rng.uniform(-1, 1, size=(12, B, 6)) * [15, 25, 30, 50, 50, 50] + [15, 25, 0, 0, 0, 0], device="cuda")
    twin = BatchedANM6Easy(B, validate_actions=False)
    twin.reset(seed=3)
    pool = twin.state.clone()
    twin.native.set_autoreset_pool(pool)
    ck = twin.state_dict()
    want = [tuple(x.clone() for x in twin.native.step(acts[t], None, chained=True)) for t in range(8)]
    want_roll = tuple(x.clone() for x in twin.native.rollout(acts[8:]))
    lib = _capi.load_library()
    grid = (B + 3) // 4
    for k in (2**32 // grid - 3, 2**32 - 4, 2**40):  # ticket wrap, ordinal wrap, far beyond both
        env = BatchedANM6Easy(B, validate_actions=False)
        env.reset(seed=3)
        env.native.set_autoreset_pool(pool)
        env.load_state_dict(ck)
        torch.cuda.synchronize()
        _capi.check(lib.anm_debug_set_launch_ordinal(env.native.h, C.c_uint64(k)), lib)
        for t in range(8):
            o, r, d = env.native.step(acts[t], None, chained=True)
            assert torch.equal(o, want[t][0]) and torch.equal(r, want[t][1]) and torch.equal(d, want[t][2]), (k, t)
        got_roll = env.native.rollout(acts[8:], chained=True)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got_roll, want_roll)), k
        assert env.native.watchdog() == [0] * 8


def test_other_solvers_same_results_subprocess():
    """Re-run the golden / oracle parity tests with the other Newton back-ends forced (ANM_SOLVER):
    the shared-memory generic kernels (default for N >= 10 buses) and the tree-elimination solver."""
    import os
    import subprocess
    import sys

    here = os.path.abspath(__file__)
    sel = "test_anm6easy_golden_trajectory or test_transition_goldens or test_batch_vs_oracle or test_radial_tree"
    # ANM_SOLVER selects the Newton back-end: generic (dense, shared memory), dense (register rows), sparse
    # (block-sparse LU, default above 9 buses); the default for the radial test networks is the tree elimination
    for val in ("generic", "dense", "sparse"):
        var = "ANM_SOLVER=" + val
        env = dict(os.environ, ANM_SOLVER=val)
        r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-k", sel, "-p", "no:cacheprovider"],
                           env=env, capture_output=True, text=True, timeout=900)  # fmt: skip
        assert r.returncode == 0, var + "\n" + r.stdout[-3000:] + r.stderr[-2000:]


def test_power_flow_identities_at_full_size():
    """BASELINE config 3 size (65536 instances): the reference's own physics identities
    (tests/simulator/test_simulator_transitions.py:189-265) on every converged instance:
    slack V = 1+0j, bus P/Q = sum of device P/Q, S = V conj(I), I = Y V, branch S/I formulas,
    sign and magnitude of the apparent power."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 65536
    env = BatchedANM6Easy(B, validate_actions=False)
    spec, cn = env.spec, env.spec.cn
    rng = np.random.default_rng(65536)
    # valid initial states without 65536 Python RNG objects: table slot + uniform Q / SoC, like init_state
    t0 = rng.integers(0, 96, B)
    s0 = np.zeros((B, 18))
    tab = spec.table
    for col, d in enumerate([1, 3, 5]):
        s0[:, d] = tab[t0, col]
        s0[:, 7 + d] = tab[t0, col] * 0.2
    for k, d in enumerate([2, 4]):
        s0[:, 15 + k] = tab[t0, 3 + k]
        s0[:, d] = tab[t0, 3 + k]
        s0[:, 7 + d] = rng.uniform(cn.devices[d].q_min, cn.devices[d].q_max, B)
    s0[:, 14] = rng.uniform(0, 1, B)
    s0[:, 17] = t0
    obs, state, conv = env.native.reset(s0, obs=env._obs, state=env.state)
    env._term_u8.copy_(1 - conv)
    full = torch.zeros(B, env.native.F, dtype=torch.float64, device=env.device)
    env._extras["full_state"] = full
    for t in range(3):
        a = torch.as_tensor(rng.uniform(spec.action_low, spec.action_high, size=(B, 6)), device=env.device)
        obs, r, term, _, _ = env.step(a)
    ok = (~term).cpu().numpy()
    assert ok.sum() > 0.9 * B and (~ok).sum() > 0
    f = full.cpu().numpy()[ok]
    sl = spec.full_state_slices()
    g = lambda k: f[:, sl[k]]  # noqa: E731
    V = g("bus_v_magn") * np.exp(1j * g("bus_v_ang"))
    I = g("bus_i_magn") * np.exp(1j * g("bus_i_ang"))  # noqa: E741
    assert np.allclose(V[:, 0], 1.0, atol=1e-12)
    # bus injections = sum of device injections
    dev_bus = np.array([cn.devices[d].bus_id for d in cn.devices])
    for b in range(6):
        assert np.allclose(g("bus_p")[:, b], g("dev_p")[:, dev_bus == b].sum(1), atol=1e-12)
        assert np.allclose(g("bus_q")[:, b], g("dev_q")[:, dev_bus == b].sum(1), atol=1e-12)
    # S = V conj(I) within the NR tolerance, I = Y V
    S = V * np.conj(I)
    assert np.abs(S.real - g("bus_p")).max() < 2e-5 and np.abs(S.imag - g("bus_q")).max() < 2e-5
    assert np.abs(I - V @ cn.Y_bus_dense.T).max() < 1e-9
    # branch flows
    for k, ((fb, tb), br) in enumerate(cn.branches.items()):
        i_from = (br.series + br.shunt) / br.tap_magn**2 * V[:, fb] - br.series / np.conj(br.tap) * V[:, tb]
        i_to = -br.series / br.tap * V[:, fb] + (br.series + br.shunt) * V[:, tb]
        s_from, s_to = V[:, fb] * np.conj(i_from), V[:, tb] * np.conj(i_to)
        assert np.allclose(g("branch_p")[:, k], s_from.real, atol=1e-10) and np.allclose(g("branch_q")[:, k], s_from.imag, atol=1e-10)
        s_app = np.sign(s_from.real) * np.maximum(np.abs(s_from), np.abs(s_to))
        assert np.allclose(g("branch_s")[:, k], s_app, atol=1e-10)
    # reward = -(clipped e_loss + clipped penalty), obs = clip(state)
    el, pe = env.e_loss.cpu().numpy()[ok], env.penalty.cpu().numpy()[ok]
    assert np.allclose(r.cpu().numpy()[ok], -(el + pe), atol=1e-12) and np.all(pe >= 0) and np.all(pe <= 100) and np.all(np.abs(el) <= 1)
    st, ob = env.state.cpu().numpy()[ok], obs.cpu().numpy()[ok]
    assert np.array_equal(ob, np.clip(st, spec.obs_low, spec.obs_high))
    # terminal rows: zeros and the terminal reward
    bad = ~ok
    assert np.all(obs.cpu().numpy()[bad] == 0) and np.all(np.isin(r.cpu().numpy()[bad], [0.0, -100 / (1 - 0.995)]))


def _tree_network(seed=5):
    """8-bus radial feeder with an off-nominal, phase-shifting transformer and a bus with 3 children."""
    N_ = None
    rng = np.random.default_rng(seed)
    bus = np.array([[0, 0, 132, 1.0, 1.0]] + [[i, 1, 33, 1.1, 0.9] for i in range(1, 8)], dtype=np.float64)
    #          0-1 (transformer), 1-2, 1-3, 1-4 (3 children), 2-5, 5-6 (depth 4), 3-7
    edges = [(0, 1, 1.03, 8.0), (1, 2, 1, 0), (1, 3, 1, 0), (4, 1, 1, 0), (2, 5, 0.98, 0), (5, 6, 1, 0), (7, 3, 1, 0)]
    branch = np.array([[f, t, rng.uniform(0.005, 0.03), rng.uniform(0.02, 0.08), rng.uniform(0, 0.02), 40, tap, sh]
                       for f, t, tap, sh in edges])
    device = np.array([
        [0, 0, 0, N_, 300, -300, 300, -300, N_, N_, N_, N_, N_, N_, N_],
        [1, 2, -1, 0.25, 0, -12, N_, N_, N_, N_, N_, N_, N_, N_, N_],
        [2, 4, -1, 0.2, 0, -8, N_, N_, N_, N_, N_, N_, N_, N_, N_],
        [3, 6, -1, 0.3, 0, -15, N_, N_, N_, N_, N_, N_, N_, N_, N_],
        [4, 7, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_],
        [5, 5, 2, N_, 20, 0, 20, -20, 14, N_, 10, -10, N_, N_, N_],
        [6, 3, 2, N_, 15, 0, 15, -15, 10, N_, 8, -8, N_, N_, N_],
        [7, 6, 3, N_, 10, -10, 10, -10, 6, -6, 5, -5, 40, 0, 0.9],
        [8, 1, 1, N_, 25, 0, 20, -20, N_, N_, N_, N_, N_, N_, N_],
    ], dtype=object)
    return {"baseMVA": 100.0, "bus": bus, "branch": branch, "device": device}


@pytest.mark.parametrize("stress", [1.0, 4.0])
def test_radial_tree_vs_oracle(stress):
    """Tree elimination (RadialNR) on a general radial feeder: transformer tap + phase shift (asymmetric Y),
    a bus with three children, depth 4; `stress` scales the ratings so that part of the batch diverges."""
    import anm_oracle
    from gym_anm_b200.env_spec import HostEnvSpec
    from gym_anm_b200.native import NativeBatch

    net = _tree_network()
    for row in net["device"][1:]:
        for c in (4, 5, 6, 7, 8, 9, 10, 11):
            if row[c] is not None:
                row[c] = row[c] * stress
    spec = HostEnvSpec(net, "state", 0, 0.25, 0.9, 100)
    cn = spec.cn
    B = 1024
    nb, cpu = NativeBatch(spec, B), anm_oracle.OracleEnv(spec, B)
    assert nb.sizes["lanes_per_env"] in (8, 16, 32)
    rng = np.random.default_rng(8)
    soc0 = rng.uniform(0, 0.4, (B, 1))
    nb.set_state(soc=soc0)
    cpu.soc[:] = soc0
    sl = spec.full_state_slices()
    n_div = 0
    for t in range(4):
        p_load = rng.uniform([cn.devices[i].p_min * 100 for i in cn.load_ids], 0, (B, cn.N_load))
        hi = np.array([cn.devices[i].p_max * 100 for i in cn.gen_ids])
        p_pot = rng.uniform(0, hi, (B, len(hi)))
        ctrl = list(cn.gen_ids) + list(cn.des_ids)
        p_set = rng.uniform([cn.devices[i].p_min * 100 for i in ctrl], [cn.devices[i].p_max * 100 for i in ctrl], (B, len(ctrl)))
        q_set = rng.uniform([cn.devices[i].q_min * 100 for i in ctrl], [cn.devices[i].q_max * 100 for i in ctrl], (B, len(ctrl)))
        full, r, e, pe, conv = nb.transition(p_load, p_pot, p_set, q_set)
        fc, rc, ec, pc, cc, nit = cpu.transition(p_load, p_pot, p_set, q_set)
        cg = conv.cpu().numpy().astype(bool)
        assert np.array_equal(cg, cc), (t, int((cg != cc).sum()))
        n_div += int((~cc).sum())
        f = full.cpu().numpy()
        for k in ("bus_p", "bus_q", "bus_v_magn", "bus_v_ang", "dev_p", "dev_q", "des_soc", "branch_p", "branch_q", "branch_s"):
            assert rel_err(f[cc][:, sl[k]], fc[cc][:, sl[k]]) < RTOL, (t, k)
        assert rel_err(r.cpu().numpy()[cc], rc[cc]) < 1e-7
    assert (n_div > 0) == (stress > 1.0)


def test_rollout_custom_next_vars_wide_action_vs_oracle():
    """Rollout and step on a custom radial network with caller-supplied next_vars and an action row wider than the
    lane group (12 entries on 8 lanes: the non-prefetched unpack path), against the C oracle and against each other."""
    import anm_oracle
    from gym_anm_b200.env_spec import HostEnvSpec
    from gym_anm_b200.native import NativeBatch

    N_ = None
    net = _tree_network()
    extra = np.array([
        [9, 4, 2, N_, 12, 0, 12, -12, 8, N_, 6, -6, N_, N_, N_],
        [10, 2, 3, N_, 8, -8, 8, -8, 5, -5, 4, -4, 30, 0, 0.92],
    ], dtype=object)
    net["device"] = np.vstack([net["device"], extra])
    spec = HostEnvSpec(net, "state", 1, 0.25, 0.9, 100, np.array([[0, 1000]]), (1, 100))
    cn = spec.cn
    B, T = 512, 24
    A = len(spec.action_low)
    assert A == 12
    nb, nb2, cpu = NativeBatch(spec, B), NativeBatch(spec, B), anm_oracle.OracleEnv(spec, B)
    assert nb.sizes["lanes_per_env"] == 8
    rng = np.random.default_rng(12)
    D, ns = cn.N_device, cn.N_des
    pos = {d: k for k, d in enumerate(cn.devices)}
    s0 = np.zeros((B, spec.state_N))
    for i in cn.load_ids:
        s0[:, pos[i]] = rng.uniform(cn.devices[i].p_min * 100, 0, B)
    for k, i in enumerate(cn.gen_ids):
        s0[:, pos[i]] = rng.uniform(0, cn.devices[i].p_max * 100, B)
        s0[:, 2 * D + ns + k] = rng.uniform(0, cn.devices[i].p_max * 100, B)
    for k, i in enumerate(cn.des_ids):
        s0[:, 2 * D + k] = rng.uniform(0, cn.devices[i].soc_max * 100, B)
    _, _, conv_c = cpu.reset(s0)
    for n in (nb, nb2):
        _, _, conv_g = n.reset(s0)
        assert np.array_equal(conv_g.cpu().numpy().astype(bool), conv_c)
    acts = rng.uniform(spec.action_low, spec.action_high, size=(T, B, A))
    nvs = np.concatenate([rng.uniform([cn.devices[i].p_min * 100 for i in cn.load_ids], 0, (T, B, cn.N_load)),
                          rng.uniform(0, [cn.devices[i].p_max * 100 for i in cn.gen_ids], (T, B, cn.N_non_slack_gen)),
                          np.broadcast_to(np.arange(1, T + 1, dtype=np.float64)[:, None, None], (T, B, 1))], axis=2)  # fmt: skip
    obs_r, rew_r, term_r = nb.rollout(acts, nvs)
    n_term = 0
    for t in range(T):
        obs_s, rew_s, term_s = nb2.step(acts[t], nvs[t])
        obs_c, r_c, term_c, _ = cpu.step(acts[t], nvs[t])
        assert torch.equal(obs_s, obs_r[t]) and torch.equal(rew_s, rew_r[t]) and torch.equal(term_s, term_r[t]), t
        assert np.array_equal(term_s.cpu().numpy().astype(bool), term_c), t
        assert rel_err(obs_s.cpu().numpy(), obs_c) < RTOL and rel_err(rew_s.cpu().numpy(), r_c) < RTOL, t
        n_term += int(term_c.sum())


def test_mpc_constant_agent_drives_batched_env_config5():
    """BASELINE config 5 in small: MPCAgentConstant(planning_steps=10, safety_margin=0.96)
    (examples/mpc_constant.py:21) picks the actions from the batched env's own state; the same action
    tensor is fed to the GPU path and the C oracle."""
    import anm_oracle
    from gym_anm_b200.agents import MPCAgentConstant
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 32
    env = BatchedANM6Easy(B, validate_actions=True)
    env.reset(seed=16384)
    agent = MPCAgentConstant(env.simulator, env.action_space, env.gamma, safety_margin=0.96, planning_steps=10)
    cpu = anm_oracle.OracleEnv(env.spec, B)
    soc, aux, term = env.native.get_state()
    cpu.soc[:], cpu.aux[:], cpu.terminated[:] = soc.cpu().numpy(), aux.cpu().numpy(), 0
    total = 0.0
    for t in range(8):
        a = agent.act(env)
        assert a.shape == (B, 6)
        obs_g, r_g, term_g, _, _ = env.step(a)
        obs_c, r_c, term_c, _ = cpu.step(a)
        assert not term_c.any() and np.array_equal(term_g.cpu().numpy(), term_c)
        assert rel_err(obs_g.cpu().numpy(), obs_c) < RTOL and rel_err(r_g.cpu().numpy(), r_c) < RTOL
        total += float(r_g.mean())
    assert total / 8 > -5.0  # the MPC policy operates the grid at low cost


def test_c_abi_error_paths_and_edge_inputs():
    """The C ABI reports errors through return codes + anm_last_error (nothing throws), refuses networks the
    reference solver cannot handle, and treats NaN / out-of-range inputs like the reference's arithmetic
    would (NaN mismatch -> not converged -> terminated)."""
    import ctypes as C

    from gym_anm_b200 import _capi
    from gym_anm_b200.anm6 import BatchedANM6Easy
    from gym_anm_b200.env_spec import HostEnvSpec, anm6easy_spec
    from gym_anm_b200.errors import BusSpecError, NativeLibraryError
    from gym_anm_b200.native import NativeBatch
    from gym_anm_b200.networks import anm6_network

    lib = _capi.load_library()
    spec = anm6easy_spec()
    net, env, keep = spec.descs()
    h = C.c_void_p()
    assert lib.anm_create(C.byref(net), C.byref(env), 8, 99, C.byref(h)) != 0 and b"device" in lib.anm_last_error()
    env.n_state = 17
    assert lib.anm_create(C.byref(net), C.byref(env), 8, 0, C.byref(h)) != 0 and b"n_state" in lib.anm_last_error()
    # slack bus must be bus 0 (solve_load_flow.py:171): refused on the host before the library is reached
    bad = anm6_network()
    bad["bus"][0, 1], bad["bus"][1, 1] = 1, 0
    bad["device"][0][1] = 1
    with pytest.raises(BusSpecError):
        NativeBatch(HostEnvSpec(bad, "state", 0, 0.25, 0.9, 100), 4)
    # a step without next_vars on an environment that has no built-in table is an error, not a crash
    nb = NativeBatch(HostEnvSpec(anm6_network(), "state", 1, 0.25, 0.9, 100, np.array([[0, 95]])), 4)
    with pytest.raises(NativeLibraryError):
        nb.step(np.zeros((4, 6)), None)
    # NaN action -> NaN injections -> NaN mismatch -> terminated, like `nan > tol == False` in the reference
    e = BatchedANM6Easy(4, validate_actions=False)
    e.reset(seed=1)
    a = np.tile(e.spec.action_high * 0.5, (4, 1))
    a[2, 0] = np.nan
    obs, r, term, _, _ = e.step(a)
    assert term.cpu().numpy().tolist() == [False, False, True, False]
    assert float(r[2]) == -100 / (1 - 0.995) and np.all(obs[2].cpu().numpy() == 0)
    # out-of-box actions are rejected by the Python layer exactly like `assert action_space.contains(action)`
    e2 = BatchedANM6Easy(2)
    e2.reset(seed=2)
    with pytest.raises(AssertionError):
        e2.step(np.tile(e2.spec.action_high * 2, (2, 1)))


def _two_envs(B, seed, pool=True):
    from gym_anm_b200.anm6 import BatchedANM6Easy

    envs = []
    for _ in range(2):
        env = BatchedANM6Easy(B, validate_actions=False)
        env.reset(seed=seed)
        if pool:
            env.native.set_autoreset_pool(env.state.clone())
        envs.append(env)
    return envs


@pytest.mark.parametrize("autoreset", [True, False])
def test_rollout_equals_stepwise(autoreset):
    """anm_rollout (one launch, every instance runs its T steps on its own) returns, slice by slice,
    bit-identical results to T separate anm_step calls -- with and without auto-reset (terminated
    instances then stay at zeros / 0.0 / True, anm_env.py:365-367)."""
    B, T = 4096, 96
    a_env, b_env = _two_envs(B, 21, pool=autoreset)
    na, nb = a_env.native, b_env.native
    rng = np.random.default_rng(5)
    acts = torch.as_tensor(rng.uniform(a_env.spec.action_low, a_env.spec.action_high, size=(T, B, 6)), device=na.device)
    obs_r, rew_r, term_r = nb.rollout(acts[:T // 2])
    obs_2, rew_2, term_2 = nb.rollout(acts[T // 2:], chained=True)  # second launch ordered per instance
    obs_r, rew_r, term_r = torch.cat([obs_r, obs_2]), torch.cat([rew_r, rew_2]), torch.cat([term_r, term_2])
    n_term = 0
    for t in range(T):
        obs, rew, term = na.step(acts[t])
        assert torch.equal(obs, obs_r[t]), t
        assert torch.equal(rew, rew_r[t]), t
        assert torch.equal(term, term_r[t]), t
        n_term += int(term.sum())
    assert n_term > 0  # divergent instances (100 Newton iterations) were in flight across launches
    for x, y in zip(na.get_state(), nb.get_state()):
        assert torch.equal(x, y)


def test_chained_steps_graph_replay_and_oracle():
    """Chained steps captured in a CUDA graph (what bench.py times) against the C oracle."""
    import anm_oracle
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B, T = 2048, 40
    env = BatchedANM6Easy(B, validate_actions=False)
    nb = env.native
    env._seed_rngs(77)
    s0 = env.init_state_batch(np.arange(B))
    cpu = anm_oracle.OracleEnv(env.spec, B)
    _, _, conv_c = cpu.reset(s0)
    _, _, conv_g = nb.reset(s0)
    assert np.array_equal(conv_g.cpu().numpy().astype(bool), conv_c)
    rng = np.random.default_rng(8)
    acts = rng.uniform(env.spec.action_low, env.spec.action_high, size=(T, B, 6))
    acts_d = torch.as_tensor(acts, device=nb.device)
    obs = nb.empty(T, B, 18)
    rew, term = nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for t in range(T):
                nb.step(acts_d[t], None, out=(obs[t], rew[t], term[t]), chained=True)
    torch.cuda.current_stream().wait_stream(side)
    soc0, aux0, term0 = nb.get_state()
    graph.replay()
    torch.cuda.synchronize()
    first = (obs.clone(), rew.clone(), term.clone())
    for t in range(T):
        obs_c, r_c, term_c, _ = cpu.step(acts[t])
        assert np.array_equal(term[t].cpu().numpy().astype(bool), term_c), t
        assert rel_err(obs[t].cpu().numpy(), obs_c) < RTOL, t
        assert rel_err(rew[t].cpu().numpy(), r_c) < RTOL, t
    # replaying from the same carried state reproduces the same bits
    nb.set_state(soc0, aux0, term0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(obs, first[0]) and torch.equal(rew, first[1]) and torch.equal(term, first[2])


def test_queued_host_steps_equal_synchronous_ones():
    """anm_step_host_async x T + anm_host_sync (zero-copy pinned buffers, chained) == anm_step_host x T;
    pageable buffers take the staged-copy path and give the same bits."""
    B, T = 1024, 48
    a_env, b_env = _two_envs(B, 33)
    na, nb = a_env.native, b_env.native
    rng = np.random.default_rng(6)
    acts = torch.as_tensor(rng.uniform(a_env.spec.action_low, a_env.spec.action_high, size=(T, B, 6)))
    acts_pin = acts.pin_memory()
    obs_q = torch.zeros(T, B, 18, dtype=torch.float64).pin_memory()
    rew_q = torch.zeros(T, B, dtype=torch.float64).pin_memory()
    term_q = torch.zeros(T, B, dtype=torch.uint8).pin_memory()
    for t in range(T):
        nb.step_host_async(acts_pin[t], None, obs_q[t], rew_q[t], term_q[t])
    nb.host_sync()
    obs_s, rew_s, term_s = np.zeros((B, 18)), np.zeros(B), np.zeros(B, dtype=np.uint8)  # pageable
    for t in range(T):
        na.step_host(acts[t].numpy(), None, obs_s, rew_s, term_s)
        assert np.array_equal(obs_s, obs_q[t].numpy()), t
        assert np.array_equal(rew_s, rew_q[t].numpy()) and np.array_equal(term_s, term_q[t].numpy()), t
    assert term_q.sum() > 0


def test_rollout_on_pinned_host_buffers():
    """anm_rollout_host_async: [T, B, .] host arrays (pinned or pageable) staged by the copy engines == device
    rollout; also with the zero-copy variant in a subprocess."""
    B, T = 512, 32
    a_env, b_env = _two_envs(B, 44)
    na, nb = a_env.native, b_env.native
    rng = np.random.default_rng(10)
    acts = torch.as_tensor(rng.uniform(a_env.spec.action_low, a_env.spec.action_high, size=(2 * T, B, 6)))
    acts_pin = acts.pin_memory()
    obs_q = torch.zeros(2 * T, B, 18, dtype=torch.float64).pin_memory()
    rew_q = torch.zeros(2 * T, B, dtype=torch.float64).pin_memory()
    term_q = torch.zeros(2 * T, B, dtype=torch.uint8).pin_memory()
    for i in range(2):
        sl = slice(i * T, (i + 1) * T)
        nb.rollout_host_async(T, acts_pin[sl], None, obs_q[sl], rew_q[sl], term_q[sl])
    nb.host_sync()
    obs_d, rew_d, term_d = na.rollout(acts.to(na.device))
    assert torch.equal(obs_d.cpu(), obs_q) and torch.equal(rew_d.cpu(), rew_q) and torch.equal(term_d.cpu(), term_q)
    import os

    if os.environ.get("ANM_HOST_ROLLOUT") == "zc":
        return  # the zero-copy variant needs pinned buffers
    # pageable arrays, same steps replayed from the same carried state
    c_env = _two_envs(B, 44)[0]
    obs_p, rew_p, term_p = np.zeros((T, B, 18)), np.zeros((T, B)), np.zeros((T, B), dtype=np.uint8)
    c_env.native.rollout_host_async(T, acts[:T].numpy(), None, obs_p, rew_p, term_p)
    c_env.native.host_sync()
    assert np.array_equal(obs_p, obs_q[:T].numpy()) and np.array_equal(term_p, term_q[:T].numpy())


def test_device_side_seeded_reset_matches_host_rng_path():
    """anm_seed + anm_reset_seeded (init_state drawn on the GPU from per-instance PCG64(SeedSequence(seed + i))
    streams, retry loop and ANM6.reset's date draw included) == the default path that draws with NumPy Generators on
    the host: bit-identical observations / states, also for later partial resets, and equal to the reference's
    reset(seed=i) goldens."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B = 1024
    host_env = BatchedANM6Easy(B, validate_actions=False)
    dev_env = BatchedANM6Easy(B, validate_actions=False, device_init=True)
    obs_h, _ = host_env.reset(seed=0)
    obs_d, _ = dev_env.reset(seed=0)
    bad = (obs_h != obs_d).any(dim=1).nonzero().flatten().tolist()
    assert not bad, (len(bad), bad[:5], obs_h[bad[0]].tolist(), obs_d[bad[0]].tolist())
    assert torch.equal(host_env.state, dev_env.state)
    for i in range(3):
        assert rel_err(obs_d[i].cpu().numpy(), load("anm6easy_traj_seed%d.npz" % i)["reset_obs"][0]) < RTOL
    rng = np.random.default_rng(4)
    for rnd in range(3):
        for t in range(12):
            a = rng.uniform(host_env.spec.action_low, host_env.spec.action_high, size=(B, 6))
            oh, rh, th, _, _ = host_env.step(a)
            od, rd, td, _, _ = dev_env.step(a)
        assert torch.equal(oh, od) and torch.equal(th, td)
        mask = th.clone()
        mask[rnd::7] = True  # the terminated ones and a few others
        oh, _ = host_env.reset(mask=mask)
        od, _ = dev_env.reset(mask=mask)
        assert torch.equal(oh, od) and torch.equal(host_env.state, dev_env.state), rnd
    assert int(th.sum()) >= 0


def test_chained_launch_may_wait_longer_than_the_watchdog_window():
    """A chained launch waits (per instance) for the previous one; the 2 s watchdog must only fire when nothing
    moves any more, not when the predecessor is simply long: two chained rollouts of ~2.7 s each."""
    from gym_anm_b200.anm6 import BatchedANM6Easy

    B, T = 8, 200_000
    env = BatchedANM6Easy(B, validate_actions=False)
    nb = env.native
    env.reset(seed=1)
    nb.set_autoreset_pool(env.state.clone())
    gen = torch.Generator(device=nb.device)
    gen.manual_seed(3)
    lo, hi = (torch.as_tensor(x, device=nb.device) for x in (env.spec.action_low, env.spec.action_high))
    acts = torch.rand((T, B, 6), dtype=torch.float64, device=nb.device, generator=gen) * (hi - lo) + lo
    out = (nb.empty(T, B, 18), nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    nb.rollout(acts, out=out)
    nb.rollout(acts, out=out, chained=True)
    # transitive waits: two chained one-step launches behind the chained rollout -- the last one waits (through its
    # predecessor) for BOTH rollouts, far beyond 2 s + 2 ms per step of its direct predecessor alone
    o1 = nb.step(acts[0], None, chained=True)
    o2 = nb.step(acts[1], None, chained=True)
    ev1.record()
    torch.cuda.synchronize()
    assert nb.watchdog()[0] == 0
    assert ev0.elapsed_time(ev1) > 3000.0  # ~2.7 s per launch today: the later ones waited beyond the 2 s window
    assert float(out[0].abs().sum()) > 0.0 and float(o2[0].abs().sum()) > 0.0 and float(o1[0].abs().sum()) > 0.0


_TIMEOUT_SCRIPT = r"""
import os, sys
import ctypes as C
sys.path.insert(0, %(root)r)
os.environ["ANM_WD_LIMIT_MS"] = "200"
import numpy as np, torch
from gym_anm_b200 import _capi
from gym_anm_b200.anm6 import BatchedANM6Easy
from gym_anm_b200.errors import NativeLibraryError
env = BatchedANM6Easy(64, validate_actions=False)
env.reset(seed=1)
a = np.random.default_rng(0).uniform(env.spec.action_low, env.spec.action_high, size=(64, 6))
env.step(a)
torch.cuda.synchronize()
lib = _capi.load_library()
_capi.check(lib.anm_debug_stall_instance(env.native.h, 17), lib)
obs, r, d, _, _ = env.step(a)          # the launch waits 0.2 s for instance 17, gives it up and ends normally
torch.cuda.synchronize()                # no launch failure: the context is alive
wd = env.native.watchdog()
assert wd[0] == 1 and wd[1] == 17, wd
try:
    env.step(a)
    raise SystemExit("the call after a time-out must fail")
except NativeLibraryError as e:
    assert "time-out" in str(e) and "rc=-5" in str(e), str(e)
x = torch.arange(8, device="cuda").double().sum()   # other CUDA work goes on
assert float(x) == 28.0
env2 = BatchedANM6Easy(64, validate_actions=False)  # and so does a new handle
env2.reset(seed=1)
env2.step(a)
torch.cuda.synchronize()
print("TIMEOUT_PATH_OK")
"""


def test_chaining_timeout_is_recoverable_subprocess():
    """The firing path of the launch-chaining watchdog: an instance whose previous launches 'never finished' makes
    the next launch time out; the kernel records the event, skips the instance and ends normally (no trap: the CUDA
    context survives), the next call on the handle fails with ANM_E_TIMEOUT, other handles and CUDA work go on."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _TIMEOUT_SCRIPT % {"root": root}], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "TIMEOUT_PATH_OK" in res.stdout, (res.stdout[-1500:], res.stderr[-1500:])


def test_chaining_off_same_results_subprocess():
    """ANM_PDL=0 (no programmatic dependent launch, every launch fully ordered), ANM_HOST_ROLLOUT=zc (queued host
    rollouts through zero-copy instead of the copy engines) and ANM_LANES=16 give the same results."""
    import os
    import subprocess
    import sys

    here = os.path.abspath(__file__)
    sel = ("test_rollout_equals_stepwise or test_chained_steps_graph_replay_and_oracle or test_queued_host_steps "
           "or test_rollout_on_pinned")
    for var in ({"ANM_PDL": "0"}, {"ANM_HOST_ROLLOUT": "zc"}, {"ANM_LANES": "16"}):
        r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-k", sel, "-p", "no:cacheprovider"],
                           env=dict(os.environ, **var), capture_output=True, text=True, timeout=900)  # fmt: skip
        assert r.returncode == 0, str(var) + "\n" + r.stdout[-3000:] + r.stderr[-2000:]
