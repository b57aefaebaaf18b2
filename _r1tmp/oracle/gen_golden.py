"""Generate the committed golden fixtures in tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the only place /root/reference
exists):   python oracle/gen_golden.py

The reference package is imported through oracle/ref_loader.py (stand-ins only for the
uninstalled cvxpy / gymnasium, see oracle/shims/).  NumPy/SciPy do the real
Newton-Raphson arithmetic (scipy.sparse.linalg.spsolve = SuperLU), so every number
stored here except the (exact) polygon projection comes from the reference's own code
path.  Fixtures:

  anm6easy_traj_seed{S}.npz  ANM6Easy-v0 trajectories under uniform-random actions
                             (reset(seed=S); actions ~ default_rng(S + 10**6)); the env is
                             reset() again (same RNG stream) after each termination.
  transitions_{net}.npz      Simulator.transition sequences on small networks (2-bus,
                             3-bus loop, 3-bus with off-nominal transformer, the
                             test_reset network, a 2-bus net with flexible devices, the
                             synthetic 30-bus feeder) incl. Simulator.reset cases.
  tables.npz                 ANM6Easy's 96-slot tables and the known-answer tables of the
                             reference's own device tests (tests/simulator/test_devices.py:
                             290-291, 556-557).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402

warnings.simplefilter("ignore")
ref_loader.load_reference()

from gym_anm.envs import ANM6Easy  # noqa: E402
from gym_anm.simulator import Simulator  # noqa: E402
from gym_anm.simulator import solve_load_flow as slf  # noqa: E402

from gym_anm_b200.networks import synth_feeder_network  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
_LAST_NR = {}
_orig_nr = slf._newton_raphson_sparse


def _spy(*a, **k):
    r = _orig_nr(*a, **k)
    _LAST_NR["n_iter"], _LAST_NR["diff"] = r[1], r[2]
    return r


slf._newton_raphson_sparse = _spy


def net_to_float(net):
    """Object arrays with None -> float arrays with NaN (np.savez cannot store None)."""
    f = lambda a: np.array([[np.nan if v is None else float(v) for v in row] for row in a], dtype=np.float64)  # noqa: E731
    return {"baseMVA": float(net["baseMVA"]), "bus": f(net["bus"]), "device": f(net["device"]), "branch": f(net["branch"])}


def full_state_pu(sim, aux=()):
    """Flatten Simulator.state (p.u. / rad) in the order of include/anm_b200.h."""
    s = sim.state
    dev_ids = list(sim.devices.keys())
    out = []
    for key, unit in (("bus_p", "pu"), ("bus_q", "pu"), ("bus_v_magn", "pu"), ("bus_v_ang", "rad"),
                      ("bus_i_magn", "pu"), ("bus_i_ang", "rad")):  # fmt: skip
        out += [s[key][unit][i] for i in sim.buses.keys()]
    out += [s["dev_p"]["pu"][i] for i in dev_ids]
    out += [s["dev_q"]["pu"][i] for i in dev_ids]
    out += [s["des_soc"]["pu"].get(i, 0.0) for i in dev_ids]
    out += [s["gen_p_max"]["pu"].get(i, 0.0) for i in dev_ids]
    for key, unit in (("branch_p", "pu"), ("branch_q", "pu"), ("branch_s", "pu"), ("branch_i_magn", "pu"),
                      ("branch_i_ang", "rad")):  # fmt: skip
        out += [s[key][unit][k] for k in sim.branches.keys()]
    out += list(aux)
    return np.array(out, dtype=np.float64)


def gen_anm6easy(seed, T):
    env = ANM6Easy()
    drawn = []
    _init = env.init_state

    def spy_init():
        s0 = _init()
        drawn.append(s0.copy())
        return s0

    env.init_state = spy_init
    rec = {k: [] for k in ("reset_s0", "reset_attempts", "actions", "obs", "state", "reward", "terminated", "e_loss", "penalty", "n_iter",
                           "full_state", "reset_before_step", "reset_obs", "reset_state", "soc_before")}  # fmt: skip
    obs, _ = env.reset(seed=seed)
    rec["reset_s0"].append(drawn[-1]), rec["reset_attempts"].append(len(drawn))
    rec["reset_before_step"].append(0)
    rec["reset_obs"].append(obs)
    rec["reset_state"].append(env.state.copy())
    rng = np.random.default_rng(seed + 10**6)
    lo, hi = env.action_space.low, env.action_space.high
    for t in range(T):
        a = rng.uniform(lo, hi)
        rec["soc_before"].append(env.simulator.devices[6].soc)
        o, r, term, trunc, info = env.step(a)
        rec["actions"].append(a)
        rec["obs"].append(o)
        rec["state"].append(env.state.copy())
        rec["reward"].append(r)
        rec["terminated"].append(term)
        rec["e_loss"].append(env.e_loss)
        rec["penalty"].append(env.penalty)
        rec["n_iter"].append(_LAST_NR["n_iter"])
        rec["full_state"].append(full_state_pu(env.simulator, aux=[float(env.state[-1])]) if not term
                                 else np.zeros(6 * 6 + 4 * 7 + 5 * 5 + 1))  # fmt: skip
        if term:
            n0 = len(drawn)
            obs, _ = env.reset()  # continues the same np_random stream
            rec["reset_s0"].append(drawn[-1]), rec["reset_attempts"].append(len(drawn) - n0)
            rec["reset_before_step"].append(t + 1)
            rec["reset_obs"].append(obs)
            rec["reset_state"].append(env.state.copy())
    out = {k: np.array(v) for k, v in rec.items()}
    out["seed"] = np.array(seed)
    out["action_low"], out["action_high"] = lo, hi
    out["obs_low"], out["obs_high"] = env.observation_space.low, env.observation_space.high
    np.savez_compressed(os.path.join(OUT, "anm6easy_traj_seed%d.npz" % seed), **out)
    print("anm6easy seed", seed, "terminations:", int(out["terminated"].sum()), "n_iter hist:",
          np.bincount(np.minimum(out["n_iter"], 101))[[2, 3, 4, 5, 6, 100]])  # fmt: skip


def _dev_groups(sim):
    from gym_anm.simulator.components import Load, Generator, StorageUnit

    loads = [i for i, d in sim.devices.items() if isinstance(d, Load)]
    gens = [i for i, d in sim.devices.items() if isinstance(d, Generator) and not d.is_slack]
    des = [i for i, d in sim.devices.items() if isinstance(d, StorageUnit)]
    return loads, gens, des


def gen_transitions(name, net, delta_t, lamb, T, seed, scale=1.0, forced=None, resets=0, cap=3.0):
    """Random Simulator.transition sequence (SoC carried from one call to the next)."""
    sim = Simulator(net, delta_t, lamb)
    loads, gens, des = _dev_groups(sim)
    rng = np.random.default_rng(seed)
    m = sim.baseMVA

    def rng_range(lo, hi):
        lo = max(lo, -cap) * m  # `cap` (p.u.) keeps the draws in the solvable range of tiny test grids
        hi = min(hi, cap) * m
        w = (hi - lo) * (scale - 1.0) / 2.0
        return rng.uniform(lo - w, hi + w)

    soc0 = []
    for i in des:
        d = sim.devices[i]
        d.soc = float(rng.uniform(d.soc_min, d.soc_max))
        soc0.append(d.soc)
    rec = {k: [] for k in ("p_load", "p_pot", "p_set", "q_set", "full_state", "reward", "e_loss", "penalty",
                           "converged", "n_iter")}  # fmt: skip
    for t in range(T):
        if forced is not None and t < len(forced):
            pl, pp, ps, qs = forced[t]
        else:
            pl = [rng_range(sim.devices[i].p_min, 0.0) for i in loads]
            pp = [rng_range(sim.devices[i].p_min, sim.devices[i].p_max) for i in gens]
            ps = [rng_range(sim.devices[i].p_min, sim.devices[i].p_max) for i in gens + des]
            qs = [rng_range(sim.devices[i].q_min, sim.devices[i].q_max) for i in gens + des]
        _, r, e, p, conv = sim.transition(dict(zip(loads, pl)), dict(zip(gens, pp)), dict(zip(gens + des, ps)),
                                          dict(zip(gens + des, qs)))  # fmt: skip
        for k, v in zip(("p_load", "p_pot", "p_set", "q_set"), (pl, pp, ps, qs)):
            rec[k].append(np.array(v, dtype=np.float64))
        rec["full_state"].append(full_state_pu(sim))
        rec["reward"].append(r), rec["e_loss"].append(e), rec["penalty"].append(p)
        rec["converged"].append(conv), rec["n_iter"].append(_LAST_NR["n_iter"])
    out = {k: np.array(v) for k, v in rec.items()}
    out["soc0"] = np.array(soc0)
    # Simulator.reset cases (s0 in MW / MVAr / MWh; simulator.py:225-293)
    n_dev, n_des, n_gen = sim.N_device, sim.N_des, sim.N_non_slack_gen
    s0s, rconv, rfull = [], [], []
    for _ in range(resets):
        s0 = np.zeros(2 * n_dev + n_des + n_gen)
        for k, (i, d) in enumerate(sim.devices.items()):
            if d.is_slack:
                continue
            s0[k] = rng_range(d.p_min, d.p_max)
            s0[n_dev + k] = rng_range(d.q_min, d.q_max)
        for k, i in enumerate(des):
            s0[2 * n_dev + k] = rng.uniform(sim.devices[i].soc_min, sim.devices[i].soc_max) * m
        for k, i in enumerate(gens):
            s0[2 * n_dev + n_des + k] = rng_range(sim.devices[i].p_min, sim.devices[i].p_max)
        conv = sim.reset(s0)
        s0s.append(s0), rconv.append(conv), rfull.append(full_state_pu(sim))
    out["reset_s0"], out["reset_converged"], out["reset_full_state"] = map(np.array, (s0s, rconv, rfull))
    out["delta_t"], out["lamb"] = np.array(delta_t), np.array(lamb)
    out["ybus_re"], out["ybus_im"] = sim.Y_bus.toarray().real, sim.Y_bus.toarray().imag
    for k, v in net_to_float(net).items():
        out["net_" + k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, "transitions_%s.npz" % name), **out)
    print(name, "T=%d converged=%d" % (T, int(out["converged"].sum())), "resets conv:", int(np.sum(rconv)), "/", resets,
          "n_iter max", out["n_iter"].max())  # fmt: skip


N_ = None


def small_networks():
    slack = [0, 0, 0, N_, 200, -200, 200, -200, N_, N_, N_, N_, N_, N_, N_]
    nets = {}
    nets["2bus"] = (
        {
            "baseMVA": 1,
            "bus": np.array([[0, 0, 50, 1.0, 1.0], [1, 1, 50, 1.1, 0.9]]),
            "branch": np.array([[0, 1, 0.01, 0.1, 0.0, 32, 1, 0]]),
            "device": np.array([slack, [1, 1, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_]]),
        },
        0.5,
    )
    dev5 = np.array(
        [
            slack,
            [1, 1, -1, 0.2, 0, -10, N_, N_, N_, N_, N_, N_, N_, N_, N_],
            [2, 1, 1, N_, 200, 0, 200, -200, N_, N_, N_, N_, N_, N_, N_],
            [3, 2, 2, N_, 200, 0, 200, -200, N_, N_, N_, N_, N_, N_, N_],
            [4, 2, 3, N_, 200, -200, 200, -200, N_, N_, N_, N_, 100, 0, 0.9],
        ]
    )
    bus3 = np.array([[0, 0, 50, 1.0, 1.0], [1, 1, 50, 1.1, 0.9], [2, 1, 50, 1.1, 0.9]])
    nets["3bus_loop"] = (
        {
            "baseMVA": 1,
            "bus": bus3,
            "branch": np.array(
                [[0, 1, 0.01, 0.1, 0.0, 30, 1, 0], [1, 2, 0.02, 0.3, 0.2, 30, 1, 0], [2, 0, 0.05, 0.2, 0.1, 30, 1, 0]]
            ),
            "device": dev5,
        },
        0.5,
    )
    nets["3bus_xfmr"] = (
        {
            "baseMVA": 1,
            "bus": bus3,
            "branch": np.array(
                [[0, 1, 0.01, 0.1, 0.0, 30, 1, 0], [1, 2, 0.02, 0.3, 0.2, 30, 1.05, 20], [2, 0, 0.05, 0.2, 0.1, 30, 1, 0]]
            ),
            "device": dev5,
        },
        0.5,
    )
    dev5b = dev5.copy()
    for r in (2, 3):
        dev5b[r][4], dev5b[r][6], dev5b[r][7] = 100, 100, -100
    dev5b[4][4], dev5b[4][5], dev5b[4][6], dev5b[4][7] = 100, -100, 100, -100
    nets["3bus_reset"] = (
        {
            "baseMVA": 10,
            "bus": bus3,
            "branch": np.array(
                [[0, 1, 0.01, 0.1, 0.0, 30, 1, 0], [1, 2, 0.02, 0.3, 0.2, 30, 1, 0], [2, 0, 0.05, 0.2, 0.1, 30, 1, 0]]
            ),
            "device": dev5b,
        },
        0.5,
    )
    # the device rows of the reference's map_pq known-answer tests (test_devices.py:286, 552)
    nets["2bus_flex"] = (
        {
            "baseMVA": 100,
            "bus": np.array([[0, 0, 50, 1.0, 1.0], [1, 1, 50, 1.1, 0.9]]),
            "branch": np.array([[0, 1, 0.01, 0.1, 0.0, 32, 1, 0]]),
            "device": np.array(
                [
                    slack,
                    [1, 1, 1, N_, 10, 1, 2, -2, 9, N_, 1, -1, N_, N_, N_],
                    [2, 1, 3, N_, 10, -11, 20, -30, 5, -6, 15, -25, 1000, 0, 1],
                ]
            ),
        },
        1.0,
    )
    return nets


def flex_forced():
    """Set-points of the reference's known-answer tables, as forced transitions."""
    gen_pts = [(-1, 0.5), (5, 5), (5, -5), (12, 0), (10, 2), (10, -2)]
    des_pts = [(8.5, 18.5), (8.5, -28.5), (-9.5, 18.5), (-9.5, -28.5)]
    forced = []
    for k in range(max(len(gen_pts), len(des_pts))):
        g = gen_pts[k % len(gen_pts)]
        d = des_pts[k % len(des_pts)]
        forced.append(([], [10.0], [g[0], d[0]], [g[1], d[1]]))
    return forced, np.array(gen_pts, dtype=np.float64), np.array(des_pts, dtype=np.float64)


def gen_tables():
    from gym_anm.envs.anm6_env.anm6_easy import _get_gen_time_series, _get_load_time_series

    _, gp, dp = flex_forced()
    np.savez_compressed(
        os.path.join(OUT, "tables.npz"),
        P_loads=_get_load_time_series(),
        P_maxs=_get_gen_time_series(),
        gen_points=gp,
        gen_mapped=np.array([(1, 0.5), (5, 2), (5, -2), (10, 0), (9.5, 1.5), (9.5, -1.5)], dtype=np.float64),
        des_points=dp,
        des_mapped=np.array([(7.5, 17.5), (7.5, -27.5), (-8.5, 17.5), (-8.5, -27.5)], dtype=np.float64),
    )


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for s in (0, 1, 2):
        gen_anm6easy(s, 300)
    nets = small_networks()
    forced, _, _ = flex_forced()
    for name, (net, dt) in nets.items():
        gen_transitions(name, net, dt, 100, T=60, seed=7, scale=1.3 if name != "2bus_flex" else 1.6,
                        forced=forced if name == "2bus_flex" else None, resets=6,
                        cap=3.0 if name == "2bus_flex" else 1.0)  # fmt: skip
    gen_transitions("anm6", __import__("gym_anm_b200.networks", fromlist=["x"]).anm6_network(), 0.25, 100, T=120,
                    seed=11, scale=1.2, resets=10)  # fmt: skip
    gen_transitions("synth30", synth_feeder_network(), 0.25, 100, T=60, seed=30, scale=1.1, resets=6)
    gen_tables()
