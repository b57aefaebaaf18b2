"""Helpers shared by the tests: golden fixtures -> network dicts / specs."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def network_from_golden(g):
    """Rebuild the network dict (NaN -> None) stored by oracle/gen_golden.py."""
    dev = np.array([[None if np.isnan(v) else float(v) for v in row] for row in g["net_device"]], dtype=object)
    for row in dev:
        for c in (0, 1, 2):
            row[c] = int(row[c])
    return {"baseMVA": float(g["net_baseMVA"]), "bus": g["net_bus"].copy(), "device": dev, "branch": g["net_branch"].copy()}


def transition_spec(g):
    from gym_anm_b200.env_spec import HostEnvSpec

    return HostEnvSpec(network_from_golden(g), "state", 0, float(g["delta_t"]), 0.9, float(g["lamb"]))


def rel_err(a, b, atol=1e-9):
    """max relative error beyond an absolute floor: (|a-b| - atol)+ / |b|.  `atol` covers
    quantities that are exactly 0 in the reference (SURVEY.md section 8c parity definition)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return float("inf")
    d = np.nan_to_num(np.abs(a - b), nan=0.0, posinf=0.0) * ~((a == b) | (np.isnan(a) & np.isnan(b)))
    return float(np.max(np.maximum(d - atol, 0.0) / np.maximum(np.abs(np.nan_to_num(b, posinf=1.0, neginf=1.0)), 1e-300)))


# the list-style observation of oracle/gen_golden.py::gen_listobs (every unit of constants.py:31-48)
LIST_OBS = [("bus_v_magn", "all", "kV"), ("bus_i_magn", [1, 2], "kA"), ("bus_v_ang", "all", "degree"),
            ("branch_i_magn", "all", "pu"), ("branch_s", "all", "MVA"), ("bus_i_ang", [1], "degree"),
            ("branch_i_ang", "all", "rad"), ("bus_p", "all", "MW"), ("bus_q", [0, 2], "pu"), ("dev_q", "all", "MVAr"),
            ("des_soc", "all", "MWh"), ("gen_p_max", "all", "pu"), ("branch_p", "all", "MW"), ("branch_q", "all", "pu"),
            ("bus_v_magn", [1], "pu"), ("bus_i_magn", "all", "pu"), ("aux", "all")]  # fmt: skip


def listobs_spec(g):
    from gym_anm_b200.env_spec import HostEnvSpec

    return HostEnvSpec(network_from_golden(g), [tuple(o) for o in LIST_OBS], 1, float(g["delta_t"]), 0.9, 50,
                       np.array([[0.0, 10.0]]), (1, 100))


def listobs_noise_mask(spec):
    """True for the observation entries that are pure solver noise in the reference itself: the current injection
    (magnitude ~1e-11 p.u., arbitrary angle) of a bus that has no device."""
    cn = spec.cn
    with_dev = {d.bus_id for d in cn.devices.values()}
    mask = []
    for key, ids, _unit in spec.obs_values:
        for i in ids:
            mask.append(key in ("bus_i_magn", "bus_i_ang") and i not in with_dev)
    return np.array(mask)
