"""Network dictionaries and fixed time series shipped with the package.

* ``anm6_network()`` -- the 6-bus / 7-device / 5-branch grid behind ``ANM6Easy-v0``
  (values of reference gym_anm/envs/anm6_env/network.py:49-82; the schema is the
  one of components/constants.py:1-19, see network_spec.py).
* ``anm6easy_tables()`` -- the deterministic 96-slot daily profiles of the three loads
  and the two renewable generators (reference anm6_easy.py:77-132).
* ``synth_feeder_network()`` -- a seeded, lightly meshed N-bus distribution feeder used
  for BASELINE.json's "custom 30-bus" configuration (the reference ships no such
  network; examples/new_env_template.py is an empty skeleton).
"""
import numpy as np

N = None  # shorthand for "not specified" in the object arrays below


def anm6_network():
    net = {"baseMVA": 100.0}
    net["bus"] = np.array(
        [[0, 0, 132, 1.0, 1.0]] + [[i, 1, 33, 1.1, 0.9] for i in range(1, 6)], dtype=np.float64
    )
    #            id bus type Q/P  Pmax Pmin  Qmax Qmin  P+  P-  Q+   Q-  SoCmax SoCmin eff
    net["device"] = np.array(
        [
            [0, 0, 0, N, 200, -200, 200, -200, N, N, N, N, N, N, N],  # slack (transmission grid)
            [1, 3, -1, 0.2, 0, -10, N, N, N, N, N, N, N, N, N],  # residential load
            [2, 3, 2, N, 30, 0, 30, -30, 20, N, 15, -15, N, N, N],  # PV
            [3, 4, -1, 0.2, 0, -30, N, N, N, N, N, N, N, N, N],  # industrial load
            [4, 4, 2, N, 50, 0, 50, -50, 35, N, 20, -20, N, N, N],  # wind
            [5, 5, -1, 0.2, 0, -30, N, N, N, N, N, N, N, N, N],  # EV charging
            [6, 5, 3, N, 50, -50, 50, -50, 30, -30, 25, -25, 100, 0, 0.9],  # storage
        ],
        dtype=object,
    )
    #            from to   r       x       b   rate tap shift
    net["branch"] = np.array(
        [
            [0, 1, 0.0036, 0.1834, 0.0, 32, 1, 0],
            [1, 2, 0.03, 0.022, 0.0, 25, 1, 0],
            [1, 3, 0.0307, 0.0621, 0.0, 18, 1, 0],
            [2, 4, 0.0303, 0.0611, 0.0, 18, 1, 0],
            [2, 5, 0.0159, 0.0502, 0.0, 18, 1, 0],
        ],
        dtype=np.float64,
    )
    return net


def _daily_profile(night, morning_ramp, day, noon_ramp, noon):
    """96 quarter-hours: night(25) ramp(7) day(13) ramp(7) noon(13) and back, night(4)."""
    r1 = np.linspace(morning_ramp[0], morning_ramp[1], 7)
    r2 = np.linspace(noon_ramp[0], noon_ramp[1], 7)
    lvl = lambda v, n: v * np.ones(n)  # noqa: E731
    out = np.concatenate(
        (lvl(night, 25), r1, lvl(day, 13), r2, lvl(noon, 13), r2[::-1], lvl(day, 13), r1[::-1], lvl(night, 4))
    )
    assert out.shape == (96,)
    return out


def anm6easy_tables():
    """Return (P_loads [3, 96], P_maxs [2, 96]) in MW."""
    p1 = _daily_profile(-1.0, (-1.5, -4.5), -5.0, (-4.625, -2.375), -2.0)
    p3 = _daily_profile(-4.0, (-4.75, -9.25), -10.0, (-11.25, -18.75), -20.0)
    p5 = _daily_profile(0.0, (-3.125, -21.875), -25.0, (-21.875, -3.125), 0.0)
    p2 = _daily_profile(0.0, (0.5, 3.5), 4.0, (7.25, 36.75), 30.0)
    p4 = _daily_profile(40.0, (36.375, 14.625), 11.0, (14.725, 36.375), 40.0)
    return np.vstack((p1, p3, p5)), np.vstack((p2, p4))


def synth_feeder_network(n_bus=30, n_load=10, n_ren=6, n_des=3, n_loops=2, seed=30):
    """A deterministic synthetic distribution feeder that passes the network validators.

    Bus 0 is the 132 kV slack, bus 1 the 33 kV substation bus fed through one branch;
    the other buses hang in a shallow random tree below bus 1, plus ``n_loops`` extra
    loop-closing branches.  Branch impedances / ratings and device ratings are drawn in
    the range of the ANM6 grid, scaled so that the network solves from a flat start.
    """
    rng = np.random.default_rng(seed)
    net = {"baseMVA": 100.0}
    net["bus"] = np.array(
        [[0, 0, 132, 1.0, 1.0]] + [[i, 1, 33, 1.1, 0.9] for i in range(1, n_bus)], dtype=np.float64
    )
    branches = [[0, 1, 0.002, 0.05, 0.0, 150, 1, 0]]
    parent = {1: 0}
    for b in range(2, n_bus):
        lo = max(1, b - 6)
        p = int(rng.integers(lo, b))
        parent[b] = p
        r = float(rng.uniform(0.008, 0.02))
        x = float(rng.uniform(0.015, 0.04))
        rate = float(rng.choice([30, 40, 60]))
        branches.append([p, b, round(r, 4), round(x, 4), 0.0, rate, 1, 0])
    existing = {(min(f, t), max(f, t)) for f, t, *_ in branches}
    tries = 0
    while n_loops > 0 and tries < 1000:
        tries += 1
        f, t = sorted(int(v) for v in rng.integers(2, n_bus, 2))
        if f == t or (f, t) in existing:
            continue
        existing.add((f, t))
        branches.append([f, t, round(float(rng.uniform(0.01, 0.03)), 4), round(float(rng.uniform(0.02, 0.06)), 4),
                         round(float(rng.uniform(0.0, 0.02)), 4), 30.0, 1, 0])  # fmt: skip
        n_loops -= 1
    net["branch"] = np.array(branches, dtype=np.float64)

    leaves = list(rng.permutation(np.arange(2, n_bus)))
    devices = [[0, 0, 0, N, 400, -400, 400, -400, N, N, N, N, N, N, N]]
    dev_id = 1
    for k in range(n_load):
        pmin = -float(rng.choice([4, 6, 8, 10, 12]))
        devices.append([dev_id, int(leaves[k % len(leaves)]), -1, 0.2, 0, pmin, N, N, N, N, N, N, N, N, N])
        dev_id += 1
    for k in range(n_ren):
        pmax = float(rng.choice([10, 15, 20, 25]))
        devices.append([dev_id, int(leaves[(n_load + k) % len(leaves)]), 2, N, pmax, 0, pmax, -pmax,
                        round(0.7 * pmax, 3), N, round(0.5 * pmax, 3), -round(0.5 * pmax, 3), N, N, N])  # fmt: skip
        dev_id += 1
    for k in range(n_des):
        pm = float(rng.choice([10, 15, 20]))
        devices.append([dev_id, int(leaves[(n_load + n_ren + k) % len(leaves)]), 3, N, pm, -pm, pm, -pm,
                        0.6 * pm, -0.6 * pm, 0.5 * pm, -0.5 * pm, 4 * pm, 0, 0.9])  # fmt: skip
        dev_id += 1
    net["device"] = np.array(devices, dtype=object)
    return net
