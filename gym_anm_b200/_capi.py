"""ctypes binding of the C ABI declared in include/anm_b200.h.

The product path has NO CPU fallback: if the CUDA library is missing, or a call fails,
`NativeLibraryError` is raised.  Build the library with `python __graft_entry__.py`
(or `make -C gym_anm_b200/csrc`).
"""
import ctypes as C
import os

import numpy as np

from .errors import NativeLibraryError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANM_B200_LIB") or os.path.join(_HERE, "lib", "libanm_b200.so")  # override: A/B builds

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class VarSpec(C.Structure):
    _fields_ = [("quantity", C.c_int32), ("index", C.c_int32), ("mul", C.c_double), ("div", C.c_double),
                ("low", C.c_double), ("high", C.c_double)]  # fmt: skip


VAR_SPEC_DTYPE = np.dtype(
    [("quantity", "<i4"), ("index", "<i4"), ("mul", "<f8"), ("div", "<f8"), ("low", "<f8"), ("high", "<f8")]
)
assert VAR_SPEC_DTYPE.itemsize == C.sizeof(VarSpec)


class NetworkDesc(C.Structure):
    _fields_ = [
        ("n_bus", C.c_int32), ("n_dev", C.c_int32), ("n_branch", C.c_int32),
        ("base_mva", C.c_double), ("delta_t", C.c_double), ("lamb", C.c_double),
        ("bus_vmin", c_double_p), ("bus_vmax", c_double_p),
        ("dev_bus", c_int32_p), ("dev_type", c_int32_p), ("dev_param", c_double_p),
        ("br_from", c_int32_p), ("br_to", c_int32_p), ("br_param", c_double_p),
        ("ybus", c_double_p),
    ]  # fmt: skip


class EnvDesc(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("gamma", C.c_double), ("clip_e_loss", C.c_double), ("clip_penalty", C.c_double),
        ("n_state", C.c_int32), ("state_vars", C.POINTER(VarSpec)),
        ("n_obs", C.c_int32), ("obs_vars", C.POINTER(VarSpec)),
        ("table_len", C.c_int32), ("table", c_double_p),
    ]  # fmt: skip


class Sizes(C.Structure):
    _fields_ = [("num_envs", C.c_int64)] + [
        (k, C.c_int32)
        for k in ("n_bus", "n_dev", "n_branch", "n_load", "n_gen", "n_des", "n_action", "n_state", "n_obs",
                  "n_next_vars", "n_full_state", "K", "lanes_per_env", "envs_per_block", "smem_bytes")
    ]  # fmt: skip


class StepExtras(C.Structure):
    _fields_ = [("state", C.c_void_p), ("e_loss", C.c_void_p), ("penalty", C.c_void_p), ("n_iter", C.c_void_p),
                ("full_state", C.c_void_p), ("solver_stats", C.c_void_p), ("flags", C.c_uint32),
                ("reserved", C.c_uint32)]  # fmt: skip


STEP_CHAINED = 1  # anm_step_extras.flags: ANM_STEP_CHAINED (include/anm_b200.h)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def make_descs(flat, base_mva, delta_t, lamb, K, gamma, clip, state_vars, obs_vars, table=None):
    """Build (NetworkDesc, EnvDesc, keepalive) from CompiledNetwork.flat() + env settings.

    `state_vars` / `obs_vars` are NumPy arrays of VAR_SPEC_DTYPE; `table` is an optional
    [T, n_load + n_gen] float64 array (built-in periodic next_vars)."""
    keep = {k: np.ascontiguousarray(v) for k, v in flat.items()}
    net = NetworkDesc()
    net.n_bus, net.n_dev, net.n_branch = len(keep["bus_vmin"]), len(keep["dev_bus"]), len(keep["br_from"])
    net.base_mva, net.delta_t, net.lamb = float(base_mva), float(delta_t), float(lamb)
    for k in ("bus_vmin", "bus_vmax", "dev_param", "br_param", "ybus"):
        setattr(net, k, _dp(keep[k]))
    for k in ("dev_bus", "dev_type", "br_from", "br_to"):
        setattr(net, k, _ip(keep[k]))
    env = EnvDesc()
    env.K, env.gamma = int(K), float(gamma)
    env.clip_e_loss, env.clip_penalty = float(clip[0]), float(clip[1])
    keep["state_vars"] = np.ascontiguousarray(state_vars, dtype=VAR_SPEC_DTYPE)
    keep["obs_vars"] = np.ascontiguousarray(obs_vars, dtype=VAR_SPEC_DTYPE)
    env.n_state, env.n_obs = len(keep["state_vars"]), len(keep["obs_vars"])
    env.state_vars = keep["state_vars"].ctypes.data_as(C.POINTER(VarSpec))
    env.obs_vars = keep["obs_vars"].ctypes.data_as(C.POINTER(VarSpec))
    if table is not None:
        keep["table"] = np.ascontiguousarray(table, dtype=np.float64)
        env.table_len, env.table = keep["table"].shape[0], _dp(keep["table"])
    else:
        env.table_len, env.table = 0, None
    return net, env, keep


_LIB = None

_PROTOS = {
    "anm_abi_version": (C.c_int, []),
    "anm_last_error": (C.c_char_p, []),
    "anm_create": (C.c_int, [C.POINTER(NetworkDesc), C.POINTER(EnvDesc), C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "anm_destroy": (C.c_int, [C.c_void_p]),
    "anm_get_sizes": (C.c_int, [C.c_void_p, C.POINTER(Sizes)]),
    "anm_reset": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p]),
    "anm_set_reset_full_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "anm_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "anm_seed_async": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "anm_reset_seeded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 3 + [C.c_void_p]),
    "anm_rng_state_bytes": (C.c_int64, [C.c_void_p]),
    "anm_get_rng": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "anm_set_rng": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "anm_debug_stall_instance": (C.c_int, [C.c_void_p, C.c_int64]),
    "anm_debug_set_launch_ordinal": (C.c_int, [C.c_void_p, C.c_uint64]),
    "anm_debug_math": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 4 + [C.c_void_p]),
    "anm_debug_fp64_peak": (C.c_int, [C.c_int, c_double_p]),
    "anm_debug_project": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int32, C.c_double, C.c_double, c_double_p]),
    "anm_debug_project_ex": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int32, C.c_uint32, C.c_int32, C.c_int32,
                                       C.c_double, C.c_double, c_double_p]),
    "anm_debug_blob": (C.c_int64, [C.POINTER(NetworkDesc), C.POINTER(EnvDesc), C.c_void_p, C.c_int64]),
    "anm_debug_rng": (C.c_int, [C.c_uint64, C.c_int32, c_int32_p, c_double_p, c_double_p, c_double_p]),
    "anm_step": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.POINTER(StepExtras), C.c_void_p]),
    "anm_rollout": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_uint32, C.c_void_p]),
    "anm_gather_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "anm_gather_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "anm_step_packed": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int32, C.POINTER(StepExtras), C.c_void_p]),
    "anm_gather_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "anm_gather_destroy": (C.c_int, [C.c_void_p]),
    "anm_set_autoreset_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "anm_transition": (C.c_int, [C.c_void_p] + [C.c_void_p] * 9 + [C.c_void_p]),
    "anm_get_state": (C.c_int, [C.c_void_p] + [C.c_void_p] * 3 + [C.c_void_p]),
    "anm_set_state": (C.c_int, [C.c_void_p] + [C.c_void_p] * 3 + [C.c_void_p]),
    "anm_step_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "anm_reset_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "anm_step_host_async": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "anm_host_sync": (C.c_int, [C.c_void_p]),
    "anm_host_sync_previous": (C.c_int, [C.c_void_p]),
    "anm_rollout_host_async": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 5),
    "anm_host_stream": (C.c_void_p, [C.c_void_p]),
    "anm_launch_count": (C.c_int64, [C.c_void_p]),
    "anm_watchdog": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)

# include/anm_lp.h: the batched LP solver behind the MPC action source (same shared library)
_LP_PROTOS = {
    "anm_lp_last_error": (C.c_char_p, []),
    "anm_lp_create": (C.c_int, [C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int32, C.c_int,
                                C.POINTER(C.c_void_p)]),
    "anm_lp_destroy": (C.c_int, [C.c_void_p]),
    "anm_lp_solve": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7 + [C.c_void_p]),
    "anm_lp_bytes": (C.c_int64, [C.c_void_p]),
    "anm_lp_kernel": (C.c_int, [C.c_void_p]),
    "anm_debug_lp_warp_state_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int64]),
    "anm_debug_lp_solve_host_warp": (C.c_int, [C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_int64, C.c_int64,
                                               C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 7 + [C.c_int32]),
    "anm_debug_lp_state_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int64]),
    "anm_debug_lp_solve_host": (C.c_int, [C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int32,
                                          C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
}
LP_EXPORTED_SYMBOLS = tuple(_LP_PROTOS)


def load_library(path=None):
    """Load libanm_b200.so (fails loudly: there is no CPU fallback)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise NativeLibraryError(
            "CUDA extension not built: %s is missing. Run `python __graft_entry__.py` (build()) first; "
            "there is no CPU fallback." % path
        )
    try:
        lib = C.CDLL(path)
    except OSError as e:  # e.g. libcudart not found
        raise NativeLibraryError("cannot load %s: %s" % (path, e)) from e
    for name, (res, args) in list(_PROTOS.items()) + list(_LP_PROTOS.items()):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.anm_abi_version() != 3:
        raise NativeLibraryError("ABI version mismatch in %s" % path)
    _LIB = lib
    return lib


def check_lp(rc, lib=None):
    if rc != 0:
        lib = lib or load_library()
        msg = lib.anm_lp_last_error()
        raise NativeLibraryError("anm_lp call failed (rc=%d): %s" % (rc, msg.decode() if msg else "?"))


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load_library()
        msg = lib.anm_last_error()
        raise NativeLibraryError("anm_b200 call failed (rc=%d): %s" % (rc, msg.decode() if msg else "?"))
