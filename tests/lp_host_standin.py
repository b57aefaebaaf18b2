"""TEST-ONLY stand-in for gym_anm_b200.lp.BatchedLP on a machine without a GPU: the same tensors on the CPU, the solve
done by the host build of the solver code (anm_debug_lp_solve_host).  It lets the CPU suite drive the agents' device
path (bounds scatter, solution checks, action extraction) end to end; the product never imports it."""
import ctypes as C

import numpy as np
import torch

from gym_anm_b200 import _capi
from gym_anm_b200.lp import BatchedLP


class HostLP(BatchedLP):
    KERNEL = "thread"  # or "warp": the warp kernel's lane phases as host loops

    def _create(self):
        self.device = torch.device("cpu")
        nbytes = (self.lib.anm_debug_lp_state_bytes(self.n, self.m, self.stride) if self.KERNEL == "thread" else
                  self.lib.anm_debug_lp_warp_state_bytes(self.n, self.m, self.B))
        self._state = np.zeros(nbytes, np.uint8)
        self._first = True

    @property
    def kernel(self):
        return self.KERNEL

    def _launch(self, restart):
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        args = (self.n, self.m, self.A_host.ctypes.data_as(_capi.c_double_p), self.c_host.ctypes.data_as(_capi.c_double_p),
                self.B, self.stride, self.max_iter, C.c_void_p(self._state.ctypes.data), int(self._first), p(self.lo),
                p(self.up), p(restart), p(self.x), p(self.obj), p(self.status), p(self.iters))
        if self.KERNEL == "thread":
            _capi.check_lp(self.lib.anm_debug_lp_solve_host(*args), self.lib)
        else:
            _capi.check_lp(self.lib.anm_debug_lp_solve_host_warp(*args, 0), self.lib)
        self._first = False

    def close(self):
        pass


class HostWarpLP(HostLP):
    KERNEL = "warp"
