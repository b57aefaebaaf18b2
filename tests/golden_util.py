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
