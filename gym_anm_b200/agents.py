"""MPC policies of gym-anm as an *action source* for the batched engine.

`MPCAgentConstant` / `MPCAgentPerfect` mirror reference gym_anm/agents/mpc.py, mpc_constant.py,
mpc_perfect.py: an N-stage DC optimal power flow over [t+1, t+N] whose first-stage set-points
become the action.  The reference states the program with CVXPY (not installable here); it is a
linear program, restated below in standard form.  Two solvers: SciPy's HiGHS (`linprog`), one LP per
environment instance on the host (optionally over a worker-process pool), and -- `device=` -- the batched
bounded dual simplex of include/anm_lp.h on the GPU (gym_anm_b200/lp.py: the programs of a batch share the
constraint matrix, every instance keeps its tableau in HBM and is warm-started from its previous basis; the
state never leaves the device).  LP optima are not unique, so this is NOT bit-parity material: the tests
check the DC-OPF constraints the reference's own test checks (tests/test_dcopf_agent.py), compare the
objective values of the two solvers, and feed the same action tensor to the GPU path and the oracle.

Per stage i (variables in p.u.): bus angles theta[n_bus], device injections P[n_dev], per storage
unit (p_ch, p_dis) >= 0, per branch an epigraph variable z >= max(0, |B_kl (theta_k - theta_l)| -
beta rate) (mpc.py:202-319).  Objective sum_i gamma^i (sum of non-renewable generator injections
incl. the slack + lamb * sum z) (mpc.py:303-312).
"""
import os

import numpy as np
from scipy.optimize import linprog
from scipy.sparse import lil_matrix

_WORKER_AGENT = None  # per worker process: the agent whose LPs it solves (see MPCAgent._pool_map)


def _worker_init(agent):
    global _WORKER_AGENT
    os.environ["OMP_NUM_THREADS"] = "1"
    _WORKER_AGENT = agent


def _worker_solve(chunk):
    Lf, Gf, soc = chunk
    return np.stack([_WORKER_AGENT.solve_one(Lf[i], Gf[i], soc[i])[0] for i in range(Lf.shape[0])])


class MPCAgent:
    """Base class; subclasses provide `forecast_batch`.  Constructor arguments as in the reference
    (mpc.py:32-50): simulator (any object with the Simulator attributes: `BatchedSimulator` works),
    action_space, gamma, safety_margin, planning_steps."""

    def __init__(self, simulator, action_space, gamma, safety_margin=0.9, planning_steps=1, workers=0, device=None,
                 refresh=64, on_lp_failure="raise"):
        """`workers` > 0: the LPs of a batch are solved by that many worker processes (one HiGHS LP per instance and
        step stays the unit of work).  0: in this process.
        `device` (e.g. "cuda:0"): `act(env)` on a batched environment whose state lives on that device solves the
        whole batch with the GPU solver and returns the actions as a CUDA tensor (`act_device`); every instance's
        basis is dropped and rebuilt every `refresh` solves (drift insurance for the warm-started tableaux).
        `on_lp_failure`: what `act_device` does with instances whose solution fails its checks twice -- "raise"
        (default: LPSolverError; there is no silent CPU path) or "host" (opt-in: those instances' programs go to the
        host LP, counted in `lp_stats["host_fallbacks"]`)."""
        if on_lp_failure not in ("raise", "host"):
            raise ValueError("on_lp_failure must be 'raise' or 'host'")
        self.on_lp_failure = on_lp_failure
        self.workers, self._pool = int(workers), None
        self.device, self.refresh, self._dev = device, int(refresh), None
        self.lp_stats = {"solves": 0, "second_solves": 0, "host_fallbacks": 0}
        self.safety_margin, self.gamma, self.planning_steps = safety_margin, gamma, planning_steps
        self.action_space = action_space
        self.baseMVA, self.lamb, self.delta_t = simulator.baseMVA, simulator.lamb, simulator.delta_t
        devs, self.bus_ids = simulator.devices, list(simulator.buses.keys())
        self.device_ids = list(devs.keys())
        self.branch_ids = list(simulator.branches.keys())
        is_gen = lambda d: d.is_gen  # noqa: E731  (DeviceSpec flags, network_spec.py)
        is_load = lambda d: d.is_load  # noqa: E731
        is_des = lambda d: d.is_storage  # noqa: E731
        is_rer = lambda d: d.is_renewable  # noqa: E731
        self.load_ids = [i for i, d in devs.items() if is_load(d)]
        self.gen_ids = [i for i, d in devs.items() if is_gen(d)]
        self.non_slack_gen_ids = [i for i, d in devs.items() if is_gen(d) and not d.is_slack]
        self.gen_rer_ids = [i for i, d in devs.items() if is_rer(d)]
        self.des_ids = [i for i, d in devs.items() if is_des(d)]
        self.slack_dev_id = [i for i, d in devs.items() if d.is_slack][0]
        self.n_bus, self.n_dev, self.n_branch = len(self.bus_ids), len(self.device_ids), len(self.branch_ids)
        self.n_des, self.n_load, self.n_gen = len(self.des_ids), len(self.load_ids), len(self.non_slack_gen_ids)
        self.bus_pos = {b: k for k, b in enumerate(self.bus_ids)}
        self.dev_pos = {d: k for k, d in enumerate(self.device_ids)}
        self.dev_to_bus = {i: d.bus_id for i, d in devs.items()}
        Y = simulator.Y_bus
        self.B_bus = np.asarray((Y.toarray() if hasattr(Y, "toarray") else Y).imag)[np.ix_(self.bus_ids, self.bus_ids)]
        self.branch_rate = [br.rate for br in simulator.branches.values()]
        g = lambda ids, a: np.array([getattr(devs[i], a) for i in ids], dtype=np.float64)  # noqa: E731
        self.P_gen_min, self.P_gen_max = g(self.non_slack_gen_ids, "p_min"), g(self.non_slack_gen_ids, "p_max")
        self.P_des_min, self.P_des_max = g(self.des_ids, "p_min"), g(self.des_ids, "p_max")
        self.soc_min, self.soc_max = g(self.des_ids, "soc_min"), g(self.des_ids, "soc_max")
        self.des_eff = g(self.des_ids, "eff")
        self._build()

    # ---- LP structure (constant): columns per stage = [theta | P_dev | p_ch | p_dis | z] -------------
    def _build(self):
        nb, nd, ns, nl, N = self.n_bus, self.n_dev, self.n_des, self.n_branch, self.planning_steps
        self.stride = S = nb + nd + 2 * ns + nl
        self.o_th, self.o_p, self.o_ch, self.o_dis, self.o_z = 0, nb, nb + nd, nb + nd + ns, nb + nd + 2 * ns
        nvar = N * S
        c = np.zeros(nvar)
        A_eq, b_eq_rows = lil_matrix((N * (nb + ns + 1), nvar)), []
        A_ub = lil_matrix((N * (2 * nl + 2 * ns), nvar))
        lb, ub = np.full(nvar, -np.inf), np.full(nvar, np.inf)
        self.eq_load_rows = []  # (stage, load index) -> fixed through bounds instead (cheaper)
        re, ru = 0, 0
        self.ub_soc_rows = []
        for i in range(N):
            o = i * S
            # objective (mpc.py:303-312)
            for g in self.gen_ids:
                if g not in self.gen_rer_ids:
                    c[o + self.o_p + self.dev_pos[g]] += self.gamma**i
            c[o + self.o_z: o + self.o_z + nl] = self.gamma**i * self.lamb
            # nodal balance: sum_branches B (theta_i - theta_j) - sum_dev P = 0 (mpc.py:231-245)
            for bi in self.bus_ids:
                for (j, k) in self.branch_ids:
                    l, m = self.bus_pos[j], self.bus_pos[k]
                    if j == bi:
                        A_eq[re, o + l] += self.B_bus[l, m]
                        A_eq[re, o + m] -= self.B_bus[l, m]
                    elif k == bi:
                        A_eq[re, o + m] += self.B_bus[m, l]
                        A_eq[re, o + l] -= self.B_bus[m, l]
                for d in self.device_ids:
                    if self.dev_to_bus[d] == bi:
                        A_eq[re, o + self.o_p + self.dev_pos[d]] -= 1.0
                re += 1
            # storage: P = p_dis - p_ch (mpc.py:285-293)
            for s, d in enumerate(self.des_ids):
                A_eq[re, o + self.o_p + self.dev_pos[d]] = 1.0
                A_eq[re, o + self.o_dis + s] = -1.0
                A_eq[re, o + self.o_ch + s] = 1.0
                re += 1
            # slack angle = 0 (the reference indexes the angle vector with the slack DEVICE's position, mpc.py:300)
            A_eq[re, o + self.o_th + self.dev_pos[self.slack_dev_id]] = 1.0
            re += 1
            # branch epigraphs: +-B (theta_k - theta_l) - z <= beta rate (mpc.py:307-310)
            for b, (j, k) in enumerate(self.branch_ids):
                l, m = self.bus_pos[j], self.bus_pos[k]
                for sgn in (1.0, -1.0):
                    A_ub[ru, o + l] = sgn * self.B_bus[l, m]
                    A_ub[ru, o + m] = -sgn * self.B_bus[l, m]
                    A_ub[ru, o + self.o_z + b] = -1.0
                    ru += 1
            # SoC window: soc_min <= soc_0 + sum_{i'<=i} (p_ch dt eff - p_dis dt / eff) <= soc_max (mpc.py:287-297)
            for s in range(ns):
                for sgn in (1.0, -1.0):
                    for ii in range(i + 1):
                        oo = ii * S
                        A_ub[ru, oo + self.o_ch + s] = sgn * self.delta_t * self.des_eff[s]
                        A_ub[ru, oo + self.o_dis + s] = -sgn * self.delta_t / self.des_eff[s]
                    self.ub_soc_rows.append((ru, s, sgn))
                    ru += 1
            # bounds
            lb[o: o + nb], ub[o: o + nb] = -np.pi, np.pi
            for k, d in enumerate(self.des_ids):
                lb[o + self.o_p + self.dev_pos[d]], ub[o + self.o_p + self.dev_pos[d]] = self.P_des_min[k], self.P_des_max[k]
            lb[o + self.o_ch: o + self.o_z + nl] = 0.0
        self.c, self.A_eq, self.A_ub = c, A_eq.tocsr(), A_ub.tocsr()
        self.b_eq = np.zeros(A_eq.shape[0])
        self.b_ub = np.zeros(A_ub.shape[0])
        r = 0
        for i in range(N):
            for b in range(nl):
                self.b_ub[r] = self.b_ub[r + 1] = self.safety_margin * self.branch_rate[b]
                r += 2
            r += 2 * ns
        np.nan_to_num(self.b_ub, copy=False, posinf=1e30)
        self.lb, self.ub = lb, ub

    # ---- one LP -------------------------------------------------------------------------------------
    def instance_arrays(self, P_load_forecast, P_gen_forecast, soc0):
        """The instance-dependent part of the full program: (lb, ub, b_ub) for forecasts [n_load, N], [n_gen, N] and the
        state of charge [n_des] (p.u.) -- the values `_update_parameters` sets in the reference (mpc.py:395-420)."""
        lb, ub, b_ub = self.lb.copy(), self.ub.copy(), self.b_ub.copy()
        S = self.stride
        for i in range(self.planning_steps):
            o = i * S + self.o_p
            for k, d in enumerate(self.load_ids):
                lb[o + self.dev_pos[d]] = ub[o + self.dev_pos[d]] = P_load_forecast[k, i]
            for k, d in enumerate(self.non_slack_gen_ids):
                lo, hi = self.P_gen_min[k], min(self.P_gen_max[k], P_gen_forecast[k, i])
                lb[o + self.dev_pos[d]], ub[o + self.dev_pos[d]] = lo, max(lo, hi)
        for row, s, sgn in self.ub_soc_rows:
            b_ub[row] = (self.soc_max[s] - soc0[s]) if sgn > 0 else (soc0[s] - self.soc_min[s])
        return lb, ub, b_ub

    def solve_one(self, P_load_forecast, P_gen_forecast, soc0):
        """P_load_forecast [n_load, N], P_gen_forecast [n_gen, N], soc0 [n_des] in p.u. -> (action, result)."""
        lb, ub, b_ub = self.instance_arrays(P_load_forecast, P_gen_forecast, soc0)
        res = linprog(self.c, A_ub=self.A_ub, b_ub=b_ub, A_eq=self.A_eq, b_eq=self.b_eq,
                      bounds=np.stack([lb, ub], axis=1), method="highs")
        if res.status != 0:
            print("OPF problem is " + res.message)
            x = np.zeros_like(self.c)
        else:
            x = res.x
        m = self.baseMVA
        P_gen = [x[self.o_p + self.dev_pos[d]] * m for d in self.non_slack_gen_ids]
        P_des = [x[self.o_p + self.dev_pos[d]] * m for d in self.des_ids]
        a = np.concatenate((P_gen, [0.0] * len(P_gen), P_des, [0.0] * len(P_des)))
        return np.clip(a, self.action_space.low, self.action_space.high), res

    # ---- batched front-end ------------------------------------------------------------------------------
    def state_to_pu(self, state):
        """Split state rows [dev_p MW | dev_q MVAr | des_soc MWh | gen_p_max MW | aux] (anm_env.py:139-145)."""
        state = np.atleast_2d(np.asarray(state, dtype=np.float64))
        D, m = self.n_dev, self.baseMVA
        p_load = state[:, [self.dev_pos[d] for d in self.load_ids]] / m
        soc = state[:, 2 * D: 2 * D + self.n_des] / m
        p_gen_max = state[:, 2 * D + self.n_des: 2 * D + self.n_des + self.n_gen] / m
        return p_load, p_gen_max, soc

    def forecast_batch(self, env, p_load, p_gen_max):
        raise NotImplementedError

    def act(self, env):
        """Actions for every instance of a (batched or single) environment: ndarray [num_envs, A]
        (a single reference-shaped env gives shape [A])."""
        st = env.state
        if self.device is not None and hasattr(st, "is_cuda") and st.is_cuda and st.ndim == 2:
            return self.act_device(env)
        st = st.detach().cpu().numpy() if hasattr(st, "detach") else np.asarray(st)
        single = st.ndim == 1
        p_load, p_gen_max, soc = self.state_to_pu(st)
        Lf, Gf = self.forecast_batch(env, p_load, p_gen_max)
        acts = self.solve_batch(Lf, Gf, soc)
        return acts[0] if single else acts

    def solve_batch(self, Lf, Gf, soc):
        """One LP per row: Lf [B, n_load, N], Gf [B, n_gen, N], soc [B, n_des] (p.u.) -> actions [B, A]."""
        B = Lf.shape[0]
        if self.workers <= 0 or B < 2 * self.workers:
            return np.stack([self.solve_one(Lf[i], Gf[i], soc[i])[0] for i in range(B)])
        if self._pool is None:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor

            clone = self.__class__.__new__(self.__class__)  # the agent without its pool (what the workers need)
            clone.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_pool"})
            clone._pool, clone.workers = None, 0
            self._pool = ProcessPoolExecutor(self.workers, mp_context=mp.get_context("spawn"), initializer=_worker_init,
                                             initargs=(clone,))
        n_chunk = min(B, self.workers * 4)
        bounds = np.linspace(0, B, n_chunk + 1).astype(int)
        chunks = [(Lf[a:b], Gf[a:b], soc[a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        return np.concatenate(list(self._pool.map(_worker_solve, chunks)), axis=0)

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        if self._dev is not None:
            self._dev.lp.close()
            self._dev = None

    # ---- the batch on the GPU (include/anm_lp.h) -----------------------------------------------------------
    def forecast_batch_device(self, env, state, p_load, p_gen_max):
        """Torch version of `forecast_batch`: [B, n_load, N], [B, n_gen, N] CUDA tensors."""
        raise NotImplementedError

    def act_device(self, env):
        """Actions [B, A] (CUDA tensor, MW / MVAr) for a batched environment whose state tensor is on the device: new
        bounds -> one warm-started solve of every instance's program -> first-stage set-points.  Every solution is
        checked on the device (status, bound violations recomputed from the constraint matrix, the angle box
        |theta| <= pi of mpc.py:298, the added big-M boxes); an instance that fails is solved again from scratch; if it
        fails again, LPSolverError (or, with on_lp_failure="host", the host LP for that instance).  One host look per
        step: the number of failed instances."""
        import torch

        from . import lp as _lp

        st = env.state
        B = st.shape[0]
        if self._dev is None or self._dev.B != B or self._dev.lp.device != st.device:
            if self._dev is not None:
                self._dev.lp.close()  # another batch size / device: the old tableaux go
            self._dev = _DeviceProgram(self, B, st.device)
        dv = self._dev
        m = self.baseMVA
        p_load, soc, p_gen_max = st[:, dv.load_cols] / m, st[:, dv.soc_cols] / m, st[:, dv.gen_cols] / m
        Lf, Gf = self.forecast_batch_device(env, st, p_load, p_gen_max)
        N, n = self.planning_steps, dv.red.n
        lp = dv.lp
        lf = Lf.reshape(B, -1).t()
        lp.lo[dv.col_load, :B] = lf
        lp.up[dv.col_load, :B] = lf
        lp.up[dv.col_gen, :B] = torch.maximum(dv.gmin, torch.minimum(dv.gmax, Gf.reshape(B, -1).t()))
        lp.lo[dv.row_soc, :B] = (dv.soc_min - soc).repeat_interleave(N, dim=1).t()
        lp.up[dv.row_soc, :B] = (dv.soc_max - soc).repeat_interleave(N, dim=1).t()
        restart = dv.all_restart if (self.refresh > 0 and lp.solves > 0 and lp.solves % self.refresh == 0) else None
        lp.solve(restart)
        self.lp_stats["solves"] += 1
        bad = dv.failed()
        n_bad = int(bad.sum())  # the step's one host look
        if n_bad:
            lp.solve(bad.to(torch.uint8))
            self.lp_stats["second_solves"] += 1
            bad = dv.failed()
            n_bad = int(bad.sum())
        x = lp.x[:, :B]
        P_gen = x[dv.act_gen].t() * m
        P_des = (dv.act_des @ x).t() * m
        a = torch.cat([P_gen, torch.zeros_like(P_gen), P_des, torch.zeros_like(P_des)], dim=1)
        a = torch.minimum(torch.maximum(a, dv.a_lo), dv.a_hi)
        if n_bad and self.on_lp_failure != "host":
            from .errors import LPSolverError

            idx = torch.nonzero(bad).flatten()[:8].tolist()
            raise LPSolverError("%d of %d programs have no usable solution after a second solve from scratch (instances %s, "
                                "status %s); build the agent with on_lp_failure='host' to hand such instances to the host LP"
                                % (n_bad, B, idx, [_lp.LP_STATUS.get(int(s_), s_) for s_ in lp.status[idx].tolist()]))
        if n_bad:  # opt-in: the host LP on the full program (angles included)
            idx = torch.nonzero(bad).flatten()
            Lh, Gh, sh = Lf[idx].cpu().numpy(), Gf[idx].cpu().numpy(), soc[idx].cpu().numpy()
            rows = np.stack([self.solve_one(Lh[k], Gh[k], sh[k])[0] for k in range(len(idx))])
            a[idx] = torch.as_tensor(rows, device=a.device)
            self.lp_stats["host_fallbacks"] += len(idx)
        return a.contiguous()


class _DeviceProgram:
    """Device-side constants of one agent's reduced program for a batch of B instances (lp.reduce_dcopf)."""

    def __init__(self, agent, B, device):
        import torch

        from . import lp as _lp

        self.B = B
        self.red = red = _lp.reduce_dcopf(agent)
        self.lp = lp = _lp.BatchedLP(red.A, red.c, B, device)
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)  # noqa: E731
        lp.lo[:] = t(red.lo)[:, None]
        lp.up[:] = t(red.up)[:, None]
        D, n, N = agent.n_dev, red.n, agent.planning_steps
        self.load_cols = t([agent.dev_pos[d] for d in agent.load_ids], torch.long)
        self.soc_cols = t(np.arange(2 * D, 2 * D + agent.n_des), torch.long)
        self.gen_cols = t(np.arange(2 * D + agent.n_des, 2 * D + agent.n_des + agent.n_gen), torch.long)
        self.col_load, self.col_gen = t(red.col_load.ravel(), torch.long), t(red.col_gen.ravel(), torch.long)
        self.row_soc = t(n + red.row_soc.ravel(), torch.long)
        self.gmin, self.gmax = t(np.repeat(agent.P_gen_min, N))[:, None], t(np.repeat(agent.P_gen_max, N))[:, None]
        lp.lo[self.col_gen] = self.gmin
        self.soc_min, self.soc_max = t(agent.soc_min)[None, :], t(agent.soc_max)[None, :]
        self.theta_map, self.free_cols = t(red.theta_map), t(red.free_cols, torch.long)
        self.act_gen = t(red.col_gen[:, 0], torch.long)
        self.act_des = t(red.A[red.row_pdes[:, 0]])
        self.a_lo, self.a_hi = t(agent.action_space.low)[None, :], t(agent.action_space.high)[None, :]
        self.all_restart = torch.ones(B, dtype=torch.uint8, device=device)
        self.big = _lp.BIG

    def failed(self):
        """[B] bool: instances whose solution of the last solve cannot be used as it is."""
        lp, B = self.lp, self.B
        x = lp.x[:, :B]
        bad = (lp.status != 0) | (lp.violation() > 1e-7)
        bad |= (self.theta_map @ x).abs().amax(dim=0) > np.pi + 1e-9
        if self.free_cols.numel():
            bad |= x[self.free_cols].abs().amax(dim=0) > 0.5 * self.big
        return bad


class MPCAgentConstant(MPCAgent):
    """Constant forecasts over the horizon (reference mpc_constant.py:21-35)."""

    def forecast_batch_device(self, env, state, p_load, p_gen_max):
        N = self.planning_steps
        return p_load[:, :, None].expand(-1, -1, N), p_gen_max[:, :, None].expand(-1, -1, N)

    def forecast_batch(self, env, p_load, p_gen_max):
        N = self.planning_steps
        return np.repeat(p_load[:, :, None], N, axis=2), np.repeat(p_gen_max[:, :, None], N, axis=2)


class MPCAgentPerfect(MPCAgent):
    """Perfect forecasts from the environment's fixed daily profiles (reference mpc_perfect.py:21-37);
    needs `env.P_loads`, `env.P_maxs` (ANM6Easy) and the aux time index in the last state entry."""

    def forecast_batch_device(self, env, state, p_load, p_gen_max):
        import torch

        if getattr(self, "_tables", None) is None or self._tables[0].device != state.device:
            self._tables = tuple(torch.as_tensor(np.asarray(a, dtype=np.float64), device=state.device) / self.baseMVA
                                 for a in (env.P_loads, env.P_maxs))
        PL, PM = self._tables
        T, N = PL.shape[1], self.planning_steps
        idx = (state[:, -1].long()[:, None] + 1 + torch.arange(N, device=state.device)[None, :]) % T
        return PL[:, idx].permute(1, 0, 2), PM[:, idx].permute(1, 0, 2)

    def forecast_batch(self, env, p_load, p_gen_max):
        st = env.state
        st = np.atleast_2d(st.detach().cpu().numpy() if hasattr(st, "detach") else np.asarray(st))
        T, N = env.P_loads.shape[1], self.planning_steps
        t0 = st[:, -1].astype(int)  # time of day of the current step; the horizon starts at t0 + 1 and wraps
        idx = (t0[:, None] + 1 + np.arange(N)[None, :]) % T
        return env.P_loads[:, idx].transpose(1, 0, 2) / self.baseMVA, env.P_maxs[:, idx].transpose(1, 0, 2) / self.baseMVA
