"""ctypes wrapper around oracle/anm_oracle.c (the CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "anm_oracle.c")
LIB = os.path.join(HERE, "_build", "libanm_oracle.so")


def build(force=False):
    """gcc -O2 -fopenmp, FMA contraction off so the arithmetic is plain IEEE fp64."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
        os.path.getmtime(SRC), os.path.getmtime(os.path.join(HERE, "..", "include", "anm_b200.h"))
    ):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-Wall", SRC, "-o", LIB, "-lm"]
    subprocess.check_call(cmd)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def project(rows, p, q):
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    out = np.zeros(2)
    lib().anm_oracle_project(C.c_int(len(rows)), _p(rows), C.c_double(p), C.c_double(q), _p(out))
    return out


class OracleEnv:
    """Batched CPU environments driven by the C restatement.  `spec` is a
    gym_anm_b200.env_spec.HostEnvSpec (host-side configuration only, no native code)."""

    def __init__(self, spec, num_envs):
        self.spec, self.B = spec, int(num_envs)
        self.net, self.env, self._keep = spec.descs()
        cn = spec.cn
        self.n_des, self.n_gen, self.n_load, self.K = cn.N_des, cn.N_non_slack_gen, cn.N_load, spec.K
        self.soc = np.zeros((self.B, self.n_des))
        self.aux = np.zeros((self.B, self.K))
        self.terminated = np.ones(self.B, dtype=np.uint8)
        self.F = spec.n_full_state

    def transition(self, p_load, p_pot, p_set, q_set):
        B = self.B
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        p_load, p_pot, p_set, q_set = f(p_load), f(p_pot), f(p_set), f(q_set)
        full = np.zeros((B, self.F - self.K))
        r, e, pe = np.zeros(B), np.zeros(B), np.zeros(B)
        conv, nit = np.zeros(B, dtype=np.uint8), np.zeros(B, dtype=np.int32)
        rc = lib().anm_oracle_transition(C.byref(self.net), C.c_int64(B), _p(self.soc), _p(p_load), _p(p_pot), _p(p_set),
                                         _p(q_set), _p(full), _p(r), _p(e), _p(pe), _p(conv), _p(nit))  # fmt: skip
        assert rc == 0, rc
        return full, r, e, pe, conv.astype(bool), nit

    def reset(self, s0, mask=None):
        B = self.B
        s0 = np.ascontiguousarray(s0, dtype=np.float64)
        obs = np.zeros((B, self.env.n_obs))
        state = np.zeros((B, self.env.n_state))
        conv = np.zeros(B, dtype=np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().anm_oracle_reset(C.byref(self.net), C.byref(self.env), C.c_int64(B), _p(s0), _p(m), _p(self.soc),
                                    _p(self.aux), _p(self.terminated), _p(obs), _p(state), _p(conv))  # fmt: skip
        assert rc == 0, rc
        return obs, state, conv.astype(bool)

    def step(self, action, next_vars=None, want_full=False):
        B = self.B
        action = np.ascontiguousarray(action, dtype=np.float64)
        nv = None if next_vars is None else np.ascontiguousarray(next_vars, dtype=np.float64)
        obs = np.zeros((B, self.env.n_obs))
        state = np.zeros((B, self.env.n_state))
        r, e, pe = np.zeros(B), np.zeros(B), np.zeros(B)
        term, nit = np.zeros(B, dtype=np.uint8), np.zeros(B, dtype=np.int32)
        full = np.zeros((B, self.F)) if want_full else None
        rc = lib().anm_oracle_step(C.byref(self.net), C.byref(self.env), C.c_int64(B), _p(self.soc), _p(self.aux),
                                   _p(self.terminated), _p(action), _p(nv), _p(obs), _p(r), _p(term), _p(state), _p(e),
                                   _p(pe), _p(nit), _p(full))  # fmt: skip
        assert rc == 0, rc
        info = {"state": state, "e_loss": e, "penalty": pe, "n_iter": nit, "full_state": full}
        return obs, r, term.astype(bool), info
