"""The ANM6Easy-v0 task on the batched engine.

* `BatchedANM6Easy(num_envs)` -- B lock-stepped ANM6Easy instances (torch CUDA tensors).
* `ANM6Easy()` -- single-environment view with the reference's exact call shapes
  (reference gym_anm/envs/anm6_env/anm6_easy.py:8-74, anm6.py:113-141): NumPy in / NumPy out,
  Python float reward, bool terminated -- existing gym-anm agents drop in.

`init_state` draws, in order, integers(0, 96), uniform(q_min, q_max) for devices 2 and 4,
uniform(soc_min, soc_max) (anm6_easy.py:25-52), and a successful reset then consumes one
integers(1, 365) draw (ANM6.reset -> random_date, anm6.py:138, anm6_env/utils.py:22), so
the per-env PCG64 streams stay aligned with the reference's.
"""
import datetime as dt
import os

import numpy as np
import torch

from .anm_env import N_INIT_STATES_MAX, BatchedANMEnv
from .errors import EnvInitializationError
from .networks import anm6_network, anm6easy_tables
from .spaces import EnvBase


class BatchedANM6Easy(BatchedANMEnv):
    """`device_init=True`: `reset()` draws the initial states on the GPU (anm_seed / anm_reset_seeded) from per-instance
    PCG64 streams that are bit-identical to the NumPy Generators of the default path -- same observations for the
    same seeds, no host round trip per retry; `np_random` then no longer reflects the streams' positions."""

    def __init__(self, num_envs=1, device=None, seed=None, validate_actions=True, env_offset=0, device_init=False,
                 track_full_state=False):
        self.device_init = bool(device_init)
        self._device_seeded = False
        self.P_loads, self.P_maxs = anm6easy_tables()
        delta_t = 0.25
        table = np.ascontiguousarray(np.vstack((self.P_loads, self.P_maxs)).T)
        super().__init__(anm6_network(), "state", 1, delta_t, 0.995, 100, np.array([[0, 24 / delta_t - 1]]), (1, 100),
                         seed, num_envs=num_envs, device=device, table=table, validate_actions=validate_actions,
                         env_offset=env_offset, track_full_state=track_full_state)  # fmt: skip

    def init_state(self):
        n_dev, n_gen, n_des = 7, 2, 1
        devs = self.spec.cn.devices
        state = np.zeros(2 * n_dev + n_des + n_gen + self.K)
        rng = self.np_random
        t_0 = rng.integers(0, int(24 / self.delta_t))
        state[-1] = t_0
        for dev_id, p_load in zip([1, 3, 5], self.P_loads):
            state[dev_id] = p_load[t_0]
            state[n_dev + dev_id] = p_load[t_0] * devs[dev_id].qp_ratio
        for idx, (dev_id, p_max) in enumerate(zip([2, 4], self.P_maxs)):
            state[2 * n_dev + n_des + idx] = p_max[t_0]
            state[dev_id] = p_max[t_0]
            state[n_dev + dev_id] = rng.uniform(devs[dev_id].q_min, devs[dev_id].q_max)  # p.u. used as MVAr (quirk kept)
        state[2 * n_dev] = rng.uniform(devs[6].soc_min, devs[6].soc_max)
        return state

    def next_vars(self, s_t):
        aux = int((s_t[-1] + 1) % (24 / self.delta_t))
        return np.array([p[aux] for p in self.P_loads] + [p[aux] for p in self.P_maxs] + [aux], dtype=np.float64)

    def _reset_on_device(self, seed, options, mask):
        """ANMEnv.reset (anm_env.py:235-311) + ANM6.reset's date draw (anm6.py:138) without leaving the GPU."""
        if seed is not None:
            self.native.seed(int(seed) + self.env_offset)
            self._device_seeded = True
        elif not self._device_seeded:
            self.native.seed(int.from_bytes(os.urandom(7), "little"))
            self._device_seeded = True
        date_draw = not (options is not None and "date_init" in options)
        conv = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        m = None if mask is None else torch.as_tensor(np.asarray(torch.as_tensor(mask).cpu(), dtype=bool), device=self.device)
        self.native.reset_seeded(m, N_INIT_STATES_MAX, date_draw, obs=self._obs, state=self.state, converged=conv)
        ok = conv.bool() if m is None else (conv.bool() | ~m)
        if not bool(ok.all()):
            raise EnvInitializationError(
                "No non-terminal state found out of %d initial states for %d environment(s) of %s"
                % (N_INIT_STATES_MAX, int((~ok).sum()), type(self).__name__)
            )
        sel = slice(None) if m is None else m
        if mask is None:
            self.timestep = 0
        self._term_u8[sel] = 0
        self.e_loss[sel] = 0.0
        self.penalty[sel] = 0.0
        return self._observe(), {}

    def reset(self, *, seed=None, options=None, mask=None):
        if self.device_init:
            return self._reset_on_device(seed, options, mask)
        obs, info = super().reset(seed=seed, options=options, mask=mask)
        # ANM6.reset draws the rendering start date from the same stream (anm6.py:138).
        idx = range(self.num_envs) if mask is None else np.flatnonzero(np.asarray(torch.as_tensor(mask).cpu(), dtype=bool))
        if not (options is not None and "date_init" in options):
            for i in idx:
                self._rngs[i].integers(1, 365)
        return obs, info


def _random_date(np_random, year):
    """anm6_env/utils.py:5-24 (the stream's integers(1, 365) draw decides the day)."""
    return dt.datetime(year, 1, 1) + dt.timedelta(days=float(np_random.integers(1, 365)))


class ANM6Easy(EnvBase):
    """Single-env, reference-shaped facade over `BatchedANM6Easy(num_envs=1)` -- the class registered as
    `ANM6Easy-v0` (a `gymnasium.Env` subclass when Gymnasium is installed).

    `simulator.state` is live (the facade tracks the full electrical state), so the reference's own
    `MPCAgentConstant` / `MPCAgentPerfect` and its renderer attach unchanged.  `render()` feeds the reference's
    renderer message format (anm6.py:46-111): pass `renderer=` an object with the `start / update / close` functions
    of gym_anm/envs/anm6_env/rendering/py/rendering.py (that module itself if gym-anm is installed; the browser UI is
    out of scope here)."""

    metadata = {"render_modes": ["human"]}

    def __init__(self, device=None, renderer=None):
        self._b = BatchedANM6Easy(1, device=device, track_full_state=True)
        self.action_space, self.observation_space = self._b.action_space, self._b.observation_space
        self.K, self.gamma, self.lamb, self.delta_t = self._b.K, self._b.gamma, self._b.lamb, self._b.delta_t
        self.costs_clipping = self._b.costs_clipping
        self.P_loads, self.P_maxs = self._b.P_loads, self._b.P_maxs
        self.simulator = self._b.simulator
        self.state_N, self.observation_N = self._b.state_N, self._b.observation_N
        self.terminated, self.timestep, self.e_loss, self.penalty = False, 0, 0.0, 0.0
        self.state = None
        # rendering bookkeeping of ANM6 (anm6.py:35-44)
        self.network_specs = self.simulator.get_rendering_specs()
        self.timestep_length = dt.timedelta(minutes=int(60 * self.delta_t))
        self.date = self.date_init = None
        self.year_count = 0
        self.skipped_frames = None
        self.render_mode = None
        self.is_rendering = False
        self._renderer = renderer
        self.http_server = self.ws_server = None

    @property
    def np_random(self):
        return self._b.np_random

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        render_mode = self.render_mode
        if seed is not None:
            self._b._seed_rngs(seed)
        elif self._b._rngs is None:
            self._b._seed_rngs(None)
        # BatchedANMEnv.reset, then ANM6.reset's date draw from the same stream (anm6.py:124-140)
        obs, info = BatchedANMEnv.reset(self._b, seed=None, options=options)
        self.render_mode = render_mode
        self.year_count = 0
        if options is not None and "date_init" in options:
            self.date_init = options["date_init"]
        else:
            self.date_init = _random_date(self._b._rngs[0], 2020)
        self.date = self.date_init
        self.terminated, self.timestep, self.e_loss, self.penalty = False, 0, 0.0, 0.0
        self.state = self._b.state[0].cpu().numpy()
        self.simulator.pfe_converged = True
        return obs[0].cpu().numpy(), info

    def step(self, action):
        action = np.asarray(action)
        assert self.action_space.contains(action), "Action %r (%s) invalid." % (action, type(action))
        was_terminated = self.terminated
        obs, r, term, trunc, info = self._b.step(action[None])
        self.timestep = self._b.timestep
        self.terminated = bool(term[0])
        self.state = self._b.state[0].cpu().numpy()
        self.e_loss, self.penalty = float(self._b.e_loss[0]), float(self._b.penalty[0])
        if not was_terminated:
            self.simulator.pfe_converged = not self.terminated
        if self.date is not None:  # anm6.py:113-122
            self.date += self.timestep_length
            self.year_count = (self.date - self.date_init).days // 365
        return obs[0].cpu().numpy(), float(r[0]), self.terminated, False, info

    # ---- rendering bridge (anm6.py:46-111, 146-239) -------------------------------------------------------
    def _rendering(self):
        if self._renderer is None:
            try:  # the reference's own module, if gym-anm happens to be installed next to this package
                from gym_anm.envs.anm6_env.rendering.py import rendering as ref_rendering

                self._renderer = ref_rendering
            except Exception as e:  # noqa: BLE001
                raise NotImplementedError(
                    "no renderer: pass ANM6Easy(renderer=...) an object with start / update / close "
                    "(gym_anm/envs/anm6_env/rendering/py/rendering.py); the browser UI is not part of this package"
                ) from e
        return self._renderer

    def render_message(self):
        """The per-frame payload of the reference's renderer (anm6.py:101-111), from instance 0's full state."""
        full_state = self.simulator.state
        return dict(
            dev_p=list(full_state["dev_p"]["MW"].values()), dev_q=list(full_state["dev_q"]["MVAr"].values()),
            branch_s=list(full_state["branch_s"]["MVA"].values()), des_soc=list(full_state["des_soc"]["MWh"].values()),
            gen_p_max=list(full_state["gen_p_max"]["MW"].values()),
            bus_v_magn=list(full_state["bus_v_magn"]["pu"].values()), costs=[self.e_loss, self.penalty],
            network_collapsed=not self.simulator.pfe_converged)  # fmt: skip

    def _init_render(self, network_specs):
        dev_type = list(network_specs["dev_type"].values())
        ps = [float(np.max(np.abs(network_specs["dev_p"][i]["MW"][:2]))) for i in network_specs["dev_p"]]
        qs = [float(np.max(np.abs(network_specs["dev_q"][i]["MVAr"][:2]))) for i in network_specs["dev_q"]]
        branch_rate = [network_specs["branch_s"][b]["MVA"][1] for b in network_specs["branch_s"]]
        bus_v_min = [network_specs["bus_v"][i]["pu"][0] for i in network_specs["bus_v"]]
        bus_v_max = [network_specs["bus_v"][i]["pu"][1] for i in network_specs["bus_v"]]
        soc_max = [network_specs["des_soc"][i]["MWh"][1] for i in network_specs["des_soc"]]
        c1 = 100 if self.costs_clipping[0] is None else self.costs_clipping[0]
        c2 = 10000 if self.costs_clipping[1] is None else self.costs_clipping[1]
        self.http_server, self.ws_server = self._rendering().start(
            type(self).__name__, dev_type, ps, qs, branch_rate, bus_v_min, bus_v_max, soc_max, (c1, c2))

    def render(self, mode="human", skip_frames=0):
        if self.render_mode is None:
            if mode not in ["human"]:
                raise NotImplementedError()
            self.render_mode = mode
            self.skipped_frames = 0
            keys = ["dev_type", "dev_p", "dev_q", "branch_s", "bus_v", "des_soc"]
            self._init_render({k: self.network_specs[k] for k in keys})
            self.render(mode=mode, skip_frames=skip_frames)
            self.is_rendering = True
        else:
            self.skipped_frames = (self.skipped_frames + 1) % (skip_frames + 1)
            if self.skipped_frames:
                return
            m = self.render_message()
            self._rendering().update(getattr(self.ws_server, "address", None), self.date, self.year_count, m["dev_p"],
                                     m["dev_q"], m["branch_s"], m["des_soc"], m["gen_p_max"], m["bus_v_magn"], m["costs"],
                                     m["network_collapsed"])  # fmt: skip

    def close(self):
        if self.is_rendering:
            try:
                self._rendering().close(self.http_server, self.ws_server)
            except AttributeError:
                pass
        self.render_mode = None
        self.is_rendering = False
