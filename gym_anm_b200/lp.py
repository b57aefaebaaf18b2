"""The MPC agents' DC-OPF as a batch of linear programs on the GPU (BASELINE config 5).

Reference: gym_anm/agents/mpc.py:161-319 builds ONE parametrised program per agent, `_solve` (mpc.py:372-393) re-solves
it every step with new forecasts and the current state of charge.  For B environment instances that is B programs with
one constraint matrix and one cost vector, different bounds -- the shape `include/anm_lp.h` solves (bounded dual
simplex, one GPU thread per program, tableaux resident in HBM, warm-started from the previous step's basis).

`reduce_dcopf(agent)` turns the full program of `agents.MPCAgent._build` (angles, every device injection, storage
charge / discharge, branch epigraphs; equality + inequality rows) into that shape: the equality rows determine the bus
angles, the slack injection and the storage injections from the remaining columns, so those are eliminated
(`x_elim = G x_keep`); the angle box |theta| <= pi (mpc.py:298) is not carried as rows (it is never active on a
sensible network) but CHECKED on every solution -- an instance that violates it counts as failed (agents.act_device:
second solve, then LPSolverError, or the host LP on the full program if the agent was built with on_lp_failure="host").
"""
import ctypes as C

import numpy as np

from . import _capi
from .errors import NativeLibraryError

BIG = 1.0e4  # box added where the cost pulls a column towards an infinite bound (p.u.; checked never to be active)
LP_STATUS = {0: "optimal", 1: "infeasible", 2: "iteration limit", 3: "dual infeasible"}


class ReducedLP:
    """min c.x, lo <= x <= up, lo_r <= A x <= up_r; `lo` / `up` hold [columns | rows] and are the instance-independent
    part -- the agent overwrites the entries listed in `col_load`, `col_gen`, `row_soc` per instance."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.m, self.n = self.A.shape


def reduce_dcopf(agent):
    """`agent`: a gym_anm_b200.agents.MPCAgent (its `_build` has run).  Returns a ReducedLP."""
    N, S = agent.planning_steps, agent.stride
    nb, nl, ns = agent.n_bus, agent.n_branch, agent.n_des
    A_eq, A_ub = agent.A_eq.toarray(), agent.A_ub.toarray()
    pos = lambda i, d: i * S + agent.o_p + agent.dev_pos[d]  # noqa: E731
    elim, keep = [], []
    col_load, col_gen = np.zeros((agent.n_load, N), np.int64), np.zeros((agent.n_gen, N), np.int64)
    for i in range(N):
        o = i * S
        elim += list(range(o, o + nb)) + [pos(i, agent.slack_dev_id)] + [pos(i, d) for d in agent.des_ids]
        for k, d in enumerate(agent.load_ids):
            col_load[k, i] = len(keep)
            keep.append(pos(i, d))
        for k, d in enumerate(agent.non_slack_gen_ids):
            col_gen[k, i] = len(keep)
            keep.append(pos(i, d))
        keep += list(range(o + agent.o_ch, o + agent.o_z + nl))  # p_ch, p_dis, z
    elim, keep = np.array(elim), np.array(keep)
    assert len(elim) == A_eq.shape[0] and len(set(elim) | set(keep)) == A_eq.shape[1]
    G = -np.linalg.solve(A_eq[:, elim], A_eq[:, keep])  # x_elim = G x_keep (b_eq = 0)
    G[np.abs(G) < 1e-13] = 0.0
    c = agent.c[keep] + G.T @ agent.c[elim]
    c[np.abs(c) < 1e-13] = 0.0
    where = {v: k for k, v in enumerate(elim)}
    rows, lo_r, up_r = [], [], []
    row_soc = np.zeros((ns, N), np.int64)
    soc_rows = {row: (s, sgn) for row, s, sgn in agent.ub_soc_rows}
    stage_of_soc = {}
    for row, s, sgn in agent.ub_soc_rows:  # rows come stage by stage, a (+, -) pair per unit
        stage_of_soc.setdefault(s, [])
        if sgn > 0:
            stage_of_soc[s].append(row)
    for r in range(A_ub.shape[0]):
        if r in soc_rows:
            s, sgn = soc_rows[r]
            if sgn < 0:
                continue  # the (-) row of the pair is the lower bound of the (+) row's activity
            row_soc[s, stage_of_soc[s].index(r)] = len(rows)
            lo_r.append(0.0), up_r.append(0.0)  # per instance: [soc_min - soc0, soc_max - soc0]
        else:
            lo_r.append(-np.inf), up_r.append(agent.b_ub[r])
        rows.append(A_ub[r, keep] + A_ub[r, elim] @ G)
    row_pdes = np.zeros((ns, N), np.int64)
    for i in range(N):
        for k, d in enumerate(agent.des_ids):  # the eliminated storage injection keeps its box as a range row
            row_pdes[k, i] = len(rows)
            rows.append(G[where[pos(i, d)]])
            lo_r.append(agent.P_des_min[k]), up_r.append(agent.P_des_max[k])
    A = np.array(rows)
    A[np.abs(A) < 1e-13] = 0.0
    lo_c, up_c = agent.lb[keep].copy(), agent.ub[keep].copy()
    for j in range(len(keep)):  # finite bound on the side the cost pulls towards (anm_lp.h)
        if c[j] <= 0 and not np.isfinite(up_c[j]):
            up_c[j] = BIG
        if c[j] >= 0 and not np.isfinite(lo_c[j]):
            lo_c[j] = -BIG
    theta_rows = np.array([where[i * S + k] for i in range(N) for k in range(nb)])
    return ReducedLP(A=np.ascontiguousarray(A), c=np.ascontiguousarray(c), lo=np.concatenate([lo_c, lo_r]),
                     up=np.concatenate([up_c, up_r]), keep=keep, elim=elim, G=G, col_load=col_load, col_gen=col_gen,
                     row_soc=row_soc, row_pdes=row_pdes, theta_map=G[theta_rows], n_full=A_eq.shape[1],
                     free_cols=np.array([j for j in range(len(keep)) if up_c[j] >= BIG or lo_c[j] <= -BIG], np.int64))


def instance_bounds(red, agent, Lf, Gf, soc):
    """NumPy: per-instance [B, n + m] lower / upper bounds from forecasts Lf [B, n_load, N], Gf [B, n_gen, N] and the
    state of charge soc [B, n_des] (p.u.) -- what `MPCAgent.solve_one` does for one instance (agents.py)."""
    B = Lf.shape[0]
    lo, up = np.tile(red.lo, (B, 1)), np.tile(red.up, (B, 1))
    lo[:, red.col_load.ravel()] = up[:, red.col_load.ravel()] = Lf.reshape(B, -1)
    gmin = np.repeat(agent.P_gen_min, agent.planning_steps)[None, :]
    gmax = np.repeat(agent.P_gen_max, agent.planning_steps)[None, :]
    lo[:, red.col_gen.ravel()] = gmin
    up[:, red.col_gen.ravel()] = np.maximum(gmin, np.minimum(gmax, Gf.reshape(B, -1)))
    n = red.n
    N = agent.planning_steps
    lo[:, n + red.row_soc.ravel()] = np.repeat(agent.soc_min[None, :] - soc, N, axis=1)
    up[:, n + red.row_soc.ravel()] = np.repeat(agent.soc_max[None, :] - soc, N, axis=1)
    return lo, up


def solve_host(red, lo, up, state=None, restart=None, max_iter=0, kernel="thread", reverse_lanes=False):
    """TEST HOOK (anm_debug_lp_solve_host / _host_warp): the solver's code compiled for the host.  lo / up:
    [B, n + m].  Returns (x [B, n], obj [B], status [B], iters [B], state) -- pass `state` back in for a warm start.
    `kernel`: "thread" (one thread per program) or "warp" (the warp kernel's lane phases as loops, optionally with
    the lanes in reverse order)."""
    lib = _capi.load_library()
    B = lo.shape[0]
    n, m = red.n, red.m
    first = state is None
    if first:
        nbytes = lib.anm_debug_lp_state_bytes(n, m, B) if kernel == "thread" else lib.anm_debug_lp_warp_state_bytes(n, m, B)
        state = np.zeros(nbytes, np.uint8)
    lo_t, up_t = np.ascontiguousarray(lo.T), np.ascontiguousarray(up.T)  # interleaved [n + m, B]
    x = np.zeros((n, B))
    obj, status, iters = np.zeros(B), np.zeros(B, np.int32), np.zeros(B, np.int32)
    rs = None if restart is None else np.ascontiguousarray(restart, np.uint8)
    p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)  # noqa: E731
    args = (n, m, red.A.ctypes.data_as(_capi.c_double_p), red.c.ctypes.data_as(_capi.c_double_p), B, B, int(max_iter),
            p(state), int(first), p(lo_t), p(up_t), p(rs), p(x), p(obj), p(status), p(iters))
    if kernel == "thread":
        _capi.check_lp(lib.anm_debug_lp_solve_host(*args), lib)
    else:
        _capi.check_lp(lib.anm_debug_lp_solve_host_warp(*args, int(bool(reverse_lanes))), lib)
    return x.T.copy(), obj, status, iters, state


class BatchedLP:
    """One `anm_lp_handle` over torch CUDA tensors (interleaved [k, stride] layout).  Fill `lo` / `up`, call `solve`."""

    def __init__(self, A, c, batch, device, max_iter=0):
        import torch

        self.lib = _capi.load_library()
        self.device = torch.device(device)
        self.A_host, self.c_host = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(c, np.float64)
        self.m, self.n = self.A_host.shape
        self.B, self.max_iter = int(batch), int(max_iter)
        self.stride = (self.B + 31) // 32 * 32
        self.h = None
        self._create()
        kw = dict(device=self.device)
        self.lo = torch.zeros((self.n + self.m, self.stride), dtype=torch.float64, **kw)
        self.up = torch.zeros((self.n + self.m, self.stride), dtype=torch.float64, **kw)
        self.x = torch.zeros((self.n, self.stride), dtype=torch.float64, **kw)
        self.obj = torch.zeros(self.B, dtype=torch.float64, **kw)
        self.status = torch.zeros(self.B, dtype=torch.int32, **kw)
        self.iters = torch.zeros(self.B, dtype=torch.int32, **kw)
        self.A_dev = torch.as_tensor(self.A_host, **kw)
        self.solves = 0

    def _create(self):
        import torch

        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise NativeLibraryError("a CUDA device is required (the batched LP solver has no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _capi.check_lp(self.lib.anm_lp_create(self.n, self.m, self.A_host.ctypes.data_as(_capi.c_double_p),
                                              self.c_host.ctypes.data_as(_capi.c_double_p), self.B, self.stride,
                                              self.max_iter, int(self.device.index), C.byref(h)), self.lib)
        self.h = h

    @property
    def bytes(self):
        return int(self.lib.anm_lp_bytes(self.h))

    @property
    def kernel(self):
        return {0: "thread", 1: "warp"}[int(self.lib.anm_lp_kernel(self.h))]

    def _launch(self, restart):
        import torch

        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        _capi.check_lp(self.lib.anm_lp_solve(self.h, p(self.lo), p(self.up), p(restart), p(self.x), p(self.obj),
                                             p(self.status), p(self.iters), st), self.lib)

    def solve(self, restart=None):
        """Enqueue one solve on the current stream.  `restart`: optional [B] uint8 tensor (non-zero = cold start)."""
        self._launch(restart)
        self.solves += 1
        return self.x

    def violation(self):
        """Largest bound violation of every instance's solution, [B] (rows recomputed from A: independent of the
        tableau, so it also measures the drift of a long warm-started sequence)."""
        import torch

        x = self.x[:, : self.B]
        act = torch.cat([x, self.A_dev @ x], dim=0)
        lo, up = self.lo[:, : self.B], self.up[:, : self.B]
        return torch.clamp(torch.maximum(lo - act, act - up), min=0.0).amax(dim=0)

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.anm_lp_destroy(h)

    __del__ = close
