"""CPU tests of the host side: network compile / validation semantics, env spec, the C-ABI
library's exported symbols, sharding + gloo all-gather (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from gym_anm_b200 import _capi, errors
from gym_anm_b200.env_spec import HostEnvSpec, anm6easy_spec
from gym_anm_b200.network_spec import CompiledNetwork
from gym_anm_b200.networks import anm6_network, synth_feeder_network

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_anm6_compile_known_values():
    cn = CompiledNetwork(anm6_network(), 0.25, 100)
    assert (cn.N_bus, cn.N_device, cn.N_branch, cn.N_load, cn.N_non_slack_gen, cn.N_des) == (6, 7, 5, 3, 2, 1)
    d2, d6 = cn.devices[2], cn.devices[6]
    assert (d2.tau_1, d2.tau_2, d2.rho_1, d2.rho_2) == pytest.approx((-1.5, 1.5, 0.6, -0.6))
    assert (d6.tau_1, d6.tau_2, d6.tau_3, d6.tau_4) == pytest.approx((-1.25, 1.25, -1.25, 1.25))
    Y = cn.Y_bus_dense
    assert Y[0, 0] == pytest.approx(0.106988 - 5.450463j, rel=1e-5) and Y[0, 1] == -Y[0, 0]
    assert np.count_nonzero(Y) == 16
    spec = anm6easy_spec()
    assert spec.state_N == 18 and len(spec.obs_specs) == 18
    np.testing.assert_array_equal(spec.action_low, [0, 0, -30, -50, -50, -50])
    np.testing.assert_array_equal(spec.action_high, [30, 50, 30, 50, 50, 50])
    np.testing.assert_allclose(spec.obs_low, [-200, -10, 0, -30, 0, -30, -50, -200, -2, -30, -6, -50, -6, -50, 0, 0, 0, 0], rtol=1e-15)
    np.testing.assert_allclose(spec.obs_high, [200, 0, 30, 0, 50, 0, 50, 200, 0, 30, 0, 50, 0, 50, 100, 30, 50, 95], rtol=1e-15)


def test_admittance_with_tap_and_shift():
    """Known answer of the reference's test_simulator_basics.py:48-67 (tap=2, 90 degree shift)."""
    net = {
        "baseMVA": 10,
        "bus": np.array([[0, 1, 50, 1.1, 0.9], [2, 1, 50, 1.1, 0.9], [1, 0, 100, 1.0, 1.0]]),
        "branch": np.array([[0, 1, 0.1, 0.2, 0.3, 20, 1, 90], [1, 2, 0.4, 0.5, 0.6, 20, 2, 0]]),
        "device": np.array([
            [1, 0, -1, 0.2, 0, -10, None, None, None, None, None, None, None, None, None],
            [0, 1, 0, None, 200, -200, 200, -200, None, None, None, None, None, None, None],
            [2, 2, 2, None, 30, 0, 30, -30, None, None, None, None, None, None, None],
            [3, 2, 3, None, 50, -50, 50, -50, None, None, None, None, 100, 0, 0.9]]),
    }  # fmt: skip
    cn = CompiledNetwork(net, 1, 100)
    y01, y01s, t01 = 1 / (0.1 + 0.2j), 0.3j / 2, np.exp(1j * np.pi / 2)
    y12, y12s = 1 / (0.4 + 0.5j), 0.6j / 2
    Y = np.array([[(y01 + y01s), -y01 / np.conj(t01), 0], [-y01 / t01, (y12 + y12s) / 4 + y01 + y01s, -y12 / 2],
                  [0, -y12 / 2, y12 + y12s]])  # fmt: skip
    np.testing.assert_allclose(cn.Y_bus_dense, Y, rtol=1e-12)
    assert list(cn.devices) == [0, 1, 2, 3] and list(cn.branches) == [(0, 1), (1, 2)]
    assert [cn.buses[i].p_min * 10 for i in (0, 1, 2)] == pytest.approx([-10, -200, -50])
    with pytest.raises(errors.BusSpecError):  # the batched solver needs slack == bus 0 (solve_load_flow.py:171)
        cn.flat()


@pytest.mark.parametrize("mutate,exc", [
    (lambda n: n.__setitem__("baseMVA", 0), errors.BaseMVAError),
    (lambda n: n["bus"].__setitem__((1, 1), 0), errors.BusSpecError),                 # two slack buses
    (lambda n: n["device"].__setitem__((1, 2), 0), errors.DeviceSpecError),           # two slack devices
    (lambda n: n["device"].__setitem__((0, 1), 3), errors.DeviceSpecError),           # slack dev not at slack bus
    (lambda n: n["branch"].__setitem__((1, slice(0, 2)), [1, 0]), errors.BranchSpecError),  # parallel branch
    (lambda n: n["branch"].__setitem__((0, 2), -0.1), errors.BranchSpecError),        # r < 0
    (lambda n: n["branch"].__setitem__((0, slice(2, 4)), [0, 0]), errors.BranchSpecError),  # r = x = 0
    (lambda n: n["branch"].__setitem__((0, 6), 0), errors.BranchSpecError),           # tap <= 0
    (lambda n: n["device"].__setitem__((1, 3), None), errors.LoadSpecError),          # load without Q/P
    (lambda n: n["device"].__setitem__((1, 4), 5), errors.LoadSpecError),             # load with Pmax > 0
    (lambda n: n["device"].__setitem__((2, 4), -1), errors.GenSpecError),             # gen Pmax < 0
    (lambda n: n["device"].__setitem__((2, 8), 40), errors.GenSpecError),             # P+ > Pmax
    (lambda n: n["device"].__setitem__((6, 12), None), errors.StorageSpecError),      # storage without SoC max
    (lambda n: n["device"].__setitem__((6, 14), 1.5), errors.StorageSpecError),       # eff > 1
    (lambda n: n["device"].__setitem__((6, 5), 10), errors.StorageSpecError),         # storage Pmin > 0
    (lambda n: n["bus"].__setitem__((2, 3), 0.5), errors.BusSpecError),               # vmax < vmin
])  # fmt: skip
def test_network_validation_errors(mutate, exc):
    net = anm6_network()
    mutate(net)
    with pytest.raises(exc):
        CompiledNetwork(net, 0.25, 100)


def test_env_arg_validation_and_obs_spec():
    net = anm6_network()
    with pytest.raises(errors.ArgsError):
        HostEnvSpec(net, "state", -1, 0.25, 0.9, 100)
    with pytest.raises(errors.ArgsError):
        HostEnvSpec(net, "state", 1, 0.25, 1.5, 100, np.array([[0, 1]]))
    with pytest.raises(errors.ArgsError):
        HostEnvSpec(net, 3.0, 0, 0.25, 0.9, 100)
    with pytest.raises(errors.ObsNotSupportedError):
        HostEnvSpec(net, [("bogus", "all")], 0, 0.25, 0.9, 100)
    with pytest.raises(errors.UnitsNotSupportedError):
        HostEnvSpec(net, [("bus_p", "all", "kV")], 0, 0.25, 0.9, 100)
    with pytest.raises(errors.ObsSpaceError):
        HostEnvSpec(net, [("dev_p", [99], "MW")], 0, 0.25, 0.9, 100)
    # list-style observation (reference tests/envs/custom_obs_space.py:33-52)
    s = HostEnvSpec(net, [("bus_p", "all", "MW"), ("dev_q", [0, 2], "pu"), ("branch_s", "all", "pu"), ("bus_v_ang", [1])],
                    0, 0.25, 0.9, 100)  # fmt: skip
    assert s.obs_values[0] == ("bus_p", [0, 1, 2, 3, 4, 5], "MW") and s.obs_values[3] == ("bus_v_ang", [1], "degree")
    assert len(s.obs_specs) == 6 + 2 + 5 + 1
    np.testing.assert_allclose(s.obs_low[:8], [-200, 0, 0, -10, -30, -80, -2, -0.3])
    assert np.all(np.isinf(s.obs_low[8:13])) and s.obs_low[13] == -180
    assert s.obs_specs["mul"][13] == 180.0 and s.obs_specs["div"][13] == np.pi


def test_synth30_passes_validators():
    cn = CompiledNetwork(synth_feeder_network(), 0.25, 100)
    assert (cn.N_bus, cn.N_load, cn.N_non_slack_gen, cn.N_des) == (30, 10, 6, 3)
    cn.flat()


def test_capi_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports each function of include/anm_b200.h."""
    hdr = open(os.path.join(ROOT, "include", "anm_b200.h")).read()
    declared = set(re.findall(r"\b(anm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.anm_abi_version() == 3
    # create must fail loudly (not fall back) without a usable device / with bad arguments
    spec = anm6easy_spec()
    net, env, keep = spec.descs()
    h = ctypes.c_void_p()
    rc = lib.anm_create(ctypes.byref(net), ctypes.byref(env), 0, 0, ctypes.byref(h))
    assert rc != 0 and lib.anm_last_error()


def _project_host(lib, rows, h, p, q, dyn_mask=None, dom=(-1, -1)):
    """dyn_mask=None: the full candidate table; else the table pruned as anm_create prunes it (rows outside dyn_mask
    are static, `dom` = (static row, dynamic row that dominates it))."""
    a = np.ascontiguousarray(rows[:, 0], dtype=np.float64)
    b = np.ascontiguousarray(rows[:, 1], dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    out = np.zeros(2)
    if dyn_mask is None:
        rc = lib.anm_debug_project(a.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                                   h.ctypes.data_as(_capi.c_double_p), len(h), float(p), float(q),
                                   out.ctypes.data_as(_capi.c_double_p))  # fmt: skip
    else:
        rc = lib.anm_debug_project_ex(a.ctypes.data_as(_capi.c_double_p), b.ctypes.data_as(_capi.c_double_p),
                                      h.ctypes.data_as(_capi.c_double_p), len(h), int(dyn_mask), int(dom[0]), int(dom[1]),
                                      float(p), float(q), out.ctypes.data_as(_capi.c_double_p))  # fmt: skip
    assert rc == 0
    return out


def test_projection_candidate_table_matches_the_exact_projection():
    """The candidate table that anm_create builds for the kernel (host code, anm_capi.cu: build_candidates), evaluated
    like the kernel evaluates it (anm_debug_project), against the definition of the exact projection
    (oracle/shims/cvxpy/_projection.py, the stand-in that passes the reference's own map_pq tests): ANM6Easy's
    generator / storage polygons with random and degenerate right-hand sides (p_pot = p_min, SoC at a bound, unused
    rows), random polygons, box clipping bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims", "cvxpy"))
    from _projection import project_onto_polygon

    lib = _capi.load_library()
    spec = anm6easy_spec()
    dp = np.asarray(spec.cn.flat()["dev_param"]).reshape(-1, 16)
    QP, PMIN, PMAX, QMIN, QMAX, T1, T2, T3, T4, R1, R2, R3, R4, SMIN, SMAX, EFF = range(16)

    def rows_gen(P):
        return np.array([[-1, 0, -P[PMIN]], [1, 0, P[PMAX]], [1, 0, P[PMAX]], [0, -1, -P[QMIN]], [0, 1, P[QMAX]],
                         [-P[T1], 1, P[R1]], [P[T2], -1, -P[R2]]], float)  # fmt: skip

    def rows_des(P):
        return np.array([[-1, 0, -P[PMIN]], [1, 0, P[PMAX]], [0, -1, -P[QMIN]], [0, 1, P[QMAX]], [-P[T1], 1, P[R1]],
                         [P[T2], -1, -P[R2]], [P[T3], -1, -P[R3]], [-P[T4], 1, P[R4]], [-1, 0, 0], [1, 0, 0]], float)  # fmt: skip

    rng = np.random.default_rng(0)
    worst = worst_pruned = 0.0
    for d, kind in ((2, "g"), (4, "g"), (6, "s")):
        P = dp[d]
        rows = rows_gen(P) if kind == "g" else rows_des(P)
        for it in range(1500):
            h = rows[:, 2].copy()
            if kind == "g":
                h[2] = rng.uniform(P[PMIN], P[PMAX]) if it % 7 else P[PMIN]  # p_pot, sometimes = p_min (a segment)
            else:
                soc = rng.uniform(P[SMIN], P[SMAX]) if it % 5 else rng.choice([P[SMIN], P[SMAX]])
                h[8] = -(soc - P[SMAX]) / (0.25 * P[EFF])
                h[9] = P[EFF] * (soc - P[SMIN]) / 0.25
            static_ok = True
            if it % 13 == 0:
                k = int(rng.integers(3, len(h)))
                h[k] = np.inf  # an unused row
                static_ok = (kind == "s" and k >= 8)  # the pruned table assumes that static rows stay what they are
            p, q = rng.uniform(-0.7, 0.7, 2)
            if it % 11 == 0:
                p, q = rng.uniform(-0.05, 0.2, 2)  # often inside
            want = project_onto_polygon(rows[:, :2], h, (p, q))
            got = _project_host(lib, rows, h, p, q)
            worst = max(worst, float(np.abs(want - got).max()))
            if static_ok:  # the pruned table of the kernel: static non-vertex intersections dropped, p <= p_max dominated
                got2 = _project_host(lib, rows, h, p, q, dyn_mask=4 if kind == "g" else 0x300,
                                     dom=(1, 2) if kind == "g" else (-1, -1))
                worst_pruned = max(worst_pruned, float(np.abs(want - got2).max()))
    assert worst < 1e-14, worst
    assert worst_pruned < 1e-14, worst_pruned
    # box clipping is exact (the reference's assertEqual tests, tests/simulator/test_devices.py:541-549)
    rows = rows_gen(dp[2])
    h = rows[:, 2].copy()
    for p, q in ((0.9, 0.0), (-0.3, 0.1), (0.1, 0.9), (0.05, -0.7)):
        want = project_onto_polygon(rows[:, :2], h, (p, q))
        assert np.array_equal(_project_host(lib, rows, h, p, q), want), (p, q)
    # random polygons around the origin, 3..10 rows
    for it in range(1500):
        R = int(rng.integers(3, 11))
        ang = np.sort(rng.uniform(0, 2 * np.pi, R))
        rows = np.stack([np.cos(ang), np.sin(ang), rng.uniform(0.1, 1.0, R)], axis=1)
        if it % 3 == 0:  # some axis-aligned rows
            rows[0, :2], rows[1, :2] = (1.0, 0.0), (0.0, -1.0)
        p, q = rng.uniform(-2, 2, 2)
        try:
            want = project_onto_polygon(rows[:, :2], rows[:, 2], (p, q))
        except ValueError:
            continue  # unbounded directions can make the candidate set empty only if the polygon is empty
        got = _project_host(lib, rows, rows[:, 2], p, q)
        assert np.abs(want - got).max() < 1e-13, (it, want, got)
    # NaN set-points stay NaN (no candidate is selected)
    assert np.isnan(_project_host(lib, rows_gen(dp[2]), rows_gen(dp[2])[:, 2], np.nan, 0.0)).all()


def test_rng_streams_match_numpy():
    """csrc/anm_rng.h (SeedSequence -> PCG64 -> integers / uniform, the draws of ANM6Easy.init_state and ANM6.reset)
    against NumPy itself, bit for bit, through the host-only entry point anm_debug_rng."""
    lib = _capi.load_library()
    kinds = np.array([0, 1, 1, 1, 0, 1, 1, 1, 0, 0, 1, 0], dtype=np.int32)
    lo = np.array([0, -0.3, -0.5, 0, 0, -0.3, -0.5, 0, 1, 0, 2.0, 5])
    hi = np.array([96, 0.3, 0.5, 1, 96, 0.3, 0.5, 1, 365, 1 << 20, 7.5, 6])
    for seed in list(range(300)) + [2020 + 4095, 2**32 - 1, 2**32, 2**32 + 7, 2**63 + 1]:
        out = np.zeros(len(kinds))
        rc = lib.anm_debug_rng(ctypes.c_uint64(seed), len(kinds), kinds.ctypes.data_as(_capi.c_int32_p),
                               lo.ctypes.data_as(_capi.c_double_p), hi.ctypes.data_as(_capi.c_double_p),
                               out.ctypes.data_as(_capi.c_double_p))  # fmt: skip
        assert rc == 0
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        want = [float(g.integers(int(a), int(b))) if k == 0 else g.uniform(a, b) for k, a, b in zip(kinds, lo, hi)]
        assert np.array_equal(out, np.array(want)), seed


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gym_anm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "anm_oracle" not in src and "anm_numpy" not in src and "ref_loader" not in src, f


def test_missing_extension_fails_loudly(tmp_path):
    with pytest.raises(errors.NativeLibraryError):
        _capi.load_library(str(tmp_path / "nope.so"))


def test_shard_slices():
    from gym_anm_b200.distributed import shard_slice

    for n, w in ((65536, 8), (10, 4), (7, 8), (4096, 1)):
        parts = [shard_slice(n, r, w) for r in range(w)]
        assert parts[0].start == 0 and parts[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
        sizes = [p.stop - p.start for p in parts]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch, torch.distributed as dist
import anm_oracle, anm_numpy
from gym_anm_b200.distributed import shard_slice, shard_counts, all_gather_rows
from gym_anm_b200.env_spec import anm6easy_spec
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, B = dist.get_rank(), 11                    # uneven shards on purpose (6 + 5)
spec = anm6easy_spec()
def rollout(lo, hi):
    env = anm_oracle.OracleEnv(spec, hi - lo)      # CPU stand-in for the per-rank CUDA batch
    s0 = np.stack([anm_numpy.anm6easy_init_state(spec, np.random.Generator(np.random.PCG64(np.random.SeedSequence(2020 + g)))) for g in range(lo, hi)])
    obs, _, _ = env.reset(s0)
    for t in range(5):
        a = np.stack([np.random.default_rng(10**6 + g * 100 + t).uniform(spec.action_low, spec.action_high) for g in range(lo, hi)])
        obs, r, term, _ = env.step(a)
    return obs
sl = shard_slice(B, rank, 2)
mine = torch.as_tensor(rollout(sl.start, sl.stop))
full = all_gather_rows(mine, counts=shard_counts(B, 2))
if rank == 0:
    want = torch.as_tensor(rollout(0, B))
    assert full.shape == want.shape and torch.equal(full, want), "shard-equivalence failed"
    print("SHARD_OK")
dist.barrier(); dist.destroy_process_group()
"""


def test_gloo_world2_shard_equivalence(tmp_path):
    """N>1 host logic on CPU: global-index seeding + all-gather == the single-process batch."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]  # fmt: skip
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


def test_gymnasium_registration_and_env_base_class():
    """reference gym_anm/__init__.py:8-11: `ANM6Easy-v0` is registered with Gymnasium and the registered class is a
    `gymnasium.Env` subclass (Gymnasium's wrappers assert that).  Gymnasium is not installed in this image, so the
    stand-in of oracle/shims/ plays its part: the registration call, the entry point string and the base class are what
    is pinned here (in a subprocess: the stand-in must be importable BEFORE the package)."""
    code = r"""
import sys, os, importlib
root = %r
sys.path[:0] = [root, os.path.join(root, "oracle", "shims")]
import gymnasium
import gym_anm_b200
from gymnasium.envs.registration import registry
entry = registry["ANM6Easy-v0"]
assert entry == "gym_anm_b200.anm6:ANM6Easy", entry
mod, cls = entry.split(":")
klass = getattr(importlib.import_module(mod), cls)
assert issubclass(klass, gymnasium.Env), klass.__mro__
from gym_anm_b200.anm_env import BatchedANMEnv
assert issubclass(BatchedANMEnv, gymnasium.Env)
from gym_anm_b200.spaces import Box
assert Box is gymnasium.spaces.Box
print("REGISTERED_OK")
""" % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "REGISTERED_OK" in res.stdout, res.stderr[-2000:]


def test_reference_arm_line_and_staged_reference(tmp_path):
    """`bench.py --impl reference` (runs on host cores only): one JSON line with the contract's keys, timed on the
    reference's own ANM6Easy (kind "reference").  And the archive that oracle/build_ref.py stages for the GPU box holds
    the reference's .py files byte for byte (the reference is imported from it there, unmodified)."""
    import json
    import zipfile

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import ref_loader

    if not os.path.isdir("/root/reference/gym_anm"):
        pytest.skip("reference tree not present (GPU box)")
    z = build_ref.build()
    with zipfile.ZipFile(z) as zf:
        names = [n for n in zf.namelist() if n.endswith(".py")]
        assert len(names) >= 25 and "gym_anm/simulator/solve_load_flow.py" in names
        for n in names:
            assert zf.read(n) == open(os.path.join("/root/reference", n), "rb").read(), n
    assert ref_loader.reference_available()
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "5",
                          "--cpu-steps", "50"], capture_output=True, text=True, timeout=600)  # fmt: skip
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["steps"] == 20 and d["warmup"] == 5 and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
