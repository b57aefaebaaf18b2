"""`NativeBatch`: a thin, typed wrapper of one `anm_handle` over torch CUDA tensors.

This is the only module that touches the C ABI at run time.  All tensors are float64 /
uint8 / int32 CUDA tensors on the handle's device; calls are enqueued on the current torch
CUDA stream and do not synchronise.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from .errors import NativeLibraryError


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _host_ptr(a):
    """Host buffer -> void*: NumPy array, CPU torch tensor, or a raw address (int) the caller computed once."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())


class NativeBatch:
    def __init__(self, spec, num_envs, device=None):
        if not torch.cuda.is_available():
            raise NativeLibraryError("a CUDA device is required (there is no CPU fallback)")
        self.lib = _capi.load_library()
        self.spec = spec
        self.B = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise NativeLibraryError("device must be a CUDA device, got %s" % self.device)
        if self.device.index is None:  # "cuda" means the CURRENT device (not GPU 0): the handle and every tensor live there
            self.device = torch.device("cuda", torch.cuda.current_device())
        net, env, keep = spec.descs()
        self._keep = keep
        h = C.c_void_p()
        _capi.check(self.lib.anm_create(C.byref(net), C.byref(env), C.c_int64(self.B), int(self.device.index), C.byref(h)),
                    self.lib)  # fmt: skip
        self.h = h
        sz = _capi.Sizes()
        _capi.check(self.lib.anm_get_sizes(self.h, C.byref(sz)), self.lib)
        self.sizes = {k: getattr(sz, k) for k, _ in sz._fields_}
        self.A, self.S, self.O = sz.n_action, sz.n_state, sz.n_obs
        self.NV, self.F, self.K = sz.n_next_vars, sz.n_full_state, sz.K
        self.n_load, self.n_gen, self.n_des = sz.n_load, sz.n_gen, sz.n_des
        self._pool = None

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.anm_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # -- helpers ---------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _on_device(self, t):
        if t is not None and isinstance(t, torch.Tensor) and t.is_cuda and t.device != self.device:
            raise ValueError("tensor on %s, but the handle lives on %s" % (t.device, self.device))
        return t

    def _f64(self, t, cols):
        self._on_device(t)
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            t = torch.as_tensor(np.asarray(t, dtype=np.float64) if not isinstance(t, torch.Tensor) else t,
                                dtype=torch.float64, device=self.device).contiguous()  # fmt: skip
        if tuple(t.shape) != (self.B, cols):
            raise ValueError("expected a tensor of shape (%d, %d), got %s" % (self.B, cols, tuple(t.shape)))
        return t

    # -- C-ABI calls -------------------------------------------------------------------
    def reset(self, s0, mask=None, obs=None, state=None, converged=None):
        s0 = self._f64(s0, self.S)
        obs = self.empty(self.B, self.O) if obs is None else obs
        state = self.empty(self.B, self.S) if state is None else state
        converged = torch.zeros(self.B, dtype=torch.uint8, device=self.device) if converged is None else converged
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _capi.check(self.lib.anm_reset(self.h, _ptr(s0), _ptr(mask), _ptr(obs), _ptr(state), _ptr(converged),
                                       self._stream()), self.lib)  # fmt: skip
        return obs, state, converged

    def set_reset_full_state(self, full):
        """Every later reset also writes the [B, F] full electrical state of the instances it resets (or None: off)."""
        self._reset_full = None if full is None else self._on_device(full)
        _capi.check(self.lib.anm_set_reset_full_state(self.h, _ptr(self._reset_full)), self.lib)

    def seed(self, seed_first):
        """Instance e gets the stream Generator(PCG64(SeedSequence(seed_first + e))), kept on the device."""
        _capi.check(self.lib.anm_seed_async(self.h, C.c_uint64(int(seed_first)), self._stream()), self.lib)

    def reset_seeded(self, mask=None, max_tries=100, date_draw=True, obs=None, state=None, converged=None):
        """ANMEnv.reset with ANM6Easy's init_state drawn on the device from the instances' own streams."""
        obs = self.empty(self.B, self.O) if obs is None else obs
        state = self.empty(self.B, self.S) if state is None else state
        converged = torch.zeros(self.B, dtype=torch.uint8, device=self.device) if converged is None else converged
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _capi.check(self.lib.anm_reset_seeded(self.h, _ptr(mask), int(max_tries), 1 if date_draw else 0, _ptr(obs),
                                              _ptr(state), _ptr(converged), self._stream()), self.lib)  # fmt: skip
        return obs, state, converged

    def step(self, action, next_vars=None, out=None, extras=None, chained=False):
        """out = (obs, reward, terminated) preallocated tensors or None; extras = dict of
        optional preallocated tensors among state / e_loss / penalty / n_iter / full_state.
        chained=True asserts that `action` / `next_vars` were not produced by work enqueued on the
        current stream after the previous call on this handle (ANM_STEP_CHAINED, include/anm_b200.h)."""
        action = self._f64(action, self.A)
        nv = None if next_vars is None else self._f64(next_vars, self.NV)
        if out is None:
            out = (self.empty(self.B, self.O), self.empty(self.B), self.empty(self.B, dtype=torch.uint8))
        obs, reward, term = out
        ex = None
        if extras or chained:
            ex = _capi.StepExtras()
            for k in ("state", "e_loss", "penalty", "n_iter", "full_state", "solver_stats"):
                t = (extras or {}).get(k)
                setattr(ex, k, None if t is None else t.data_ptr())
            ex.flags = _capi.STEP_CHAINED if chained else 0
        _capi.check(self.lib.anm_step(self.h, _ptr(action), _ptr(nv), _ptr(obs), _ptr(reward), _ptr(term),
                                      None if ex is None else C.byref(ex), self._stream()), self.lib)  # fmt: skip
        return obs, reward, term

    def rollout(self, actions, next_vars=None, out=None, chained=False):
        """T open-loop steps in one kernel launch: actions [T, B, A] (next_vars [T, B, NV]) -> obs [T, B, O],
        reward [T, B], terminated [T, B]; slice t equals the t-th step() (anm_rollout)."""
        ok = lambda t, c: (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()  # noqa: E731
                           and t.ndim == 3 and tuple(t.shape[1:]) == (self.B, c))
        if not ok(actions, self.A):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float64) if not isinstance(actions, torch.Tensor) else actions,
                                      dtype=torch.float64, device=self.device).contiguous()  # fmt: skip
        if not ok(actions, self.A):
            raise ValueError("expected actions of shape (T, %d, %d), got %s" % (self.B, self.A, tuple(actions.shape)))
        T = actions.shape[0]
        if next_vars is not None:
            next_vars = torch.as_tensor(next_vars, dtype=torch.float64, device=self.device).contiguous()
            if tuple(next_vars.shape) != (T, self.B, self.NV):
                raise ValueError("expected next_vars of shape (%d, %d, %d)" % (T, self.B, self.NV))
        if out is None:
            out = (self.empty(T, self.B, self.O), self.empty(T, self.B), self.empty(T, self.B, dtype=torch.uint8))
        obs, reward, term = out
        _capi.check(self.lib.anm_rollout(self.h, C.c_int64(T), _ptr(actions), _ptr(next_vars), _ptr(obs), _ptr(reward),
                                         _ptr(term), _capi.STEP_CHAINED if chained else 0, self._stream()),
                    self.lib)  # fmt: skip
        return obs, reward, term

    def transition(self, p_load, p_pot, p_set, q_set):
        n_ctrl = self.n_gen + self.n_des
        p_load, p_pot = self._f64(p_load, self.n_load), self._f64(p_pot, self.n_gen)
        p_set, q_set = self._f64(p_set, n_ctrl), self._f64(q_set, n_ctrl)
        full = self.empty(self.B, self.F - self.K)
        r, e, pe = self.empty(self.B), self.empty(self.B), self.empty(self.B)
        conv = self.empty(self.B, dtype=torch.uint8)
        _capi.check(self.lib.anm_transition(self.h, _ptr(p_load), _ptr(p_pot), _ptr(p_set), _ptr(q_set), _ptr(full),
                                            _ptr(r), _ptr(e), _ptr(pe), _ptr(conv), self._stream()), self.lib)  # fmt: skip
        return full, r, e, pe, conv

    def get_state(self):
        soc, aux = self.empty(self.B, self.n_des), self.empty(self.B, self.K)
        term = self.empty(self.B, dtype=torch.uint8)
        _capi.check(self.lib.anm_get_state(self.h, _ptr(soc), _ptr(aux), _ptr(term), self._stream()), self.lib)
        return soc, aux, term

    def set_state(self, soc=None, aux=None, terminated=None):
        f = lambda t, c: None if t is None else self._f64(t, c)  # noqa: E731
        soc, aux = f(soc, self.n_des), f(aux, self.K)
        if terminated is not None:
            terminated = torch.as_tensor(terminated, device=self.device).to(torch.uint8).contiguous()
        _capi.check(self.lib.anm_set_state(self.h, _ptr(soc), _ptr(aux), _ptr(terminated), self._stream()), self.lib)

    def get_rng(self):
        """Opaque uint8 tensor holding the device-side PCG64 streams (after seed()); see set_rng."""
        n = int(self.lib.anm_rng_state_bytes(self.h))
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        _capi.check(self.lib.anm_get_rng(self.h, _ptr(out), self._stream()), self.lib)
        return out

    def set_rng(self, blob):
        blob = torch.as_tensor(blob, dtype=torch.uint8, device=self.device).contiguous()
        if blob.numel() != int(self.lib.anm_rng_state_bytes(self.h)):
            raise ValueError("rng state of %d bytes, expected %d" % (blob.numel(), self.lib.anm_rng_state_bytes(self.h)))
        _capi.check(self.lib.anm_set_rng(self.h, _ptr(blob), self._stream()), self.lib)

    def set_autoreset_pool(self, pool):
        if pool is None:
            self._pool = None
            _capi.check(self.lib.anm_set_autoreset_pool(self.h, None, 0), self.lib)
            return
        pool = torch.as_tensor(pool, dtype=torch.float64, device=self.device).contiguous()
        assert pool.ndim == 2 and pool.shape[1] == self.S
        self._pool = pool  # keep alive: the library does not copy it
        _capi.check(self.lib.anm_set_autoreset_pool(self.h, _ptr(pool), pool.shape[0]), self.lib)

    def step_host(self, action, next_vars, obs, reward, terminated):
        """NumPy / pinned-host path: H2D + step + D2H inside the library, synchronous."""
        p = _host_ptr
        _capi.check(self.lib.anm_step_host(self.h, p(action), p(next_vars), p(obs), p(reward), p(terminated)), self.lib)

    def step_host_async(self, action, next_vars, obs, reward, terminated):
        """Queued step_host: returns at once; outputs are valid after host_sync().  One set of output buffers per
        queued step."""
        p = _host_ptr
        _capi.check(self.lib.anm_step_host_async(self.h, p(action), p(next_vars), p(obs), p(reward), p(terminated)),
                    self.lib)  # fmt: skip

    def rollout_host_async(self, T, actions, next_vars, obs, reward, terminated):
        """Queued rollout on pinned host arrays [T, B, .] (zero-copy); valid after host_sync()."""
        p = _host_ptr
        _capi.check(self.lib.anm_rollout_host_async(self.h, C.c_int64(T), p(actions), p(next_vars), p(obs), p(reward),
                                                    p(terminated)), self.lib)  # fmt: skip

    def host_sync_previous(self):
        """Wait for every queued rollout except the most recent one (consume call i-1 while call i runs)."""
        _capi.check(self.lib.anm_host_sync_previous(self.h), self.lib)

    def host_sync(self):
        _capi.check(self.lib.anm_host_sync(self.h), self.lib)

    def reset_host(self, s0, mask, obs, state, converged):
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())  # noqa: E731
        _capi.check(self.lib.anm_reset_host(self.h, p(s0), p(mask), p(obs), p(state), p(converged)), self.lib)

    @property
    def host_stream(self):
        """torch view of the library's own stream (the one step_host / reset_host run on)."""
        return torch.cuda.ExternalStream(int(self.lib.anm_host_stream(self.h)), device=self.device)

    def watchdog(self):
        """Record left by a chained launch that timed out (all zeros normally); readable after a launch failure."""
        out = (C.c_uint32 * 8)()
        _capi.check(self.lib.anm_watchdog(self.h, out), self.lib)
        return list(out)

    @property
    def launch_count(self):
        return int(self.lib.anm_launch_count(self.h))
