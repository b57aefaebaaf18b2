"""`BatchedANMEnv`: the ANMEnv surface over B lock-stepped environment instances.

Mirrors reference gym_anm/envs/anm_env.py (class ANMEnv): same constructor arguments,
`init_state()` / `next_vars(s_t)` hooks, `reset(seed=, options=)`, `step(action)`,
`observation_space` / `action_space`, attributes `state`, `terminated`, `timestep`,
`e_loss`, `penalty`, `simulator`.  The per-step body (anm_env.py:365-453) is one CUDA
kernel launch through the C ABI; tensors are row-major [num_envs, F] float64 torch CUDA
tensors.  With ``num_envs=1`` and NumPy inputs the outputs have the reference's shapes
(obs ndarray[O], float, bool, False, {}), see `SingleEnvView` in anm6.py.
"""
import warnings

import numpy as np
import torch

from .env_spec import HostEnvSpec
from .errors import EnvInitializationError, EnvNextVarsError
from .native import NativeBatch
from .simulator import BatchedSimulator
from .spaces import Box, EnvBase

N_INIT_STATES_MAX = 100  # anm_env.py:268


def _make_rng(seed):
    """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class BatchedANMEnv(EnvBase):
    metadata = {"render_modes": []}

    def __init__(self, network, observation, K, delta_t, gamma, lamb, aux_bounds=None, costs_clipping=None, seed=None,
                 *, num_envs=1, device=None, table=None, validate_actions=True, env_offset=0,
                 track_full_state=False):  # fmt: skip
        self.spec = HostEnvSpec(network, observation, K, delta_t, gamma, lamb, aux_bounds, costs_clipping, table=table)
        self.num_envs = int(num_envs)
        self.env_offset = int(env_offset)  # global index of local env 0 (multi-GPU sharding)
        self.K, self.gamma, self.lamb, self.delta_t, self.aux_bounds = K, gamma, lamb, delta_t, aux_bounds
        self.costs_clipping = self.spec.costs_clipping
        self.native = NativeBatch(self.spec, self.num_envs, device)
        self.device = self.native.device
        self.simulator = BatchedSimulator(self.spec, self.native)
        self.state_values, self.state_N = self.spec.state_values, self.spec.state_N
        self.obs_values = self.spec.obs_values
        self.action_space = Box(low=self.spec.action_low, high=self.spec.action_high, dtype=np.float64)
        if self.obs_values is not None:
            self.observation_space = Box(low=self.spec.obs_low, high=self.spec.obs_high, dtype=np.float64)
            self.observation_N = self.observation_space.shape[0]
        else:
            self.observation_space, self.observation_N = None, None
        self._act_lo = torch.as_tensor(self.spec.action_low, device=self.device)
        self._act_hi = torch.as_tensor(self.spec.action_high, device=self.device)
        self.validate_actions = validate_actions
        B, nb = self.num_envs, self.native
        # carried / output tensors (allocated once)
        self.state = torch.zeros(B, nb.S, dtype=torch.float64, device=self.device)
        self._obs = torch.zeros(B, nb.O, dtype=torch.float64, device=self.device)
        self.reward = torch.zeros(B, dtype=torch.float64, device=self.device)
        self._term_u8 = torch.ones(B, dtype=torch.uint8, device=self.device)
        self.e_loss = torch.zeros(B, dtype=torch.float64, device=self.device)
        self.penalty = torch.zeros(B, dtype=torch.float64, device=self.device)
        self.n_iter = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._extras = {"state": self.state, "e_loss": self.e_loss, "penalty": self.penalty, "n_iter": self.n_iter}
        # `simulator.state` (the nested dict the reference's renderer and MPC agents read, anm6.py:101-109, mpc.py:419)
        # needs every electrical quantity of the step: an extra [B, F] output row per instance, off by default
        self.track_full_state = bool(track_full_state)
        if self.track_full_state:
            self._full = torch.zeros(B, nb.F, dtype=torch.float64, device=self.device)
            self._extras["full_state"] = self._full
            self.native.set_reset_full_state(self._full)
            self.simulator._env = self
        self.timestep = 0
        self.render_mode = None
        self._seed = seed
        self._rngs = None
        self._env_index = 0
        if seed is not None:
            self._seed_rngs(seed)

    # ---- hooks to override (same contract as the reference, one env at a time) --------------
    def init_state(self):
        """Sample an initial state s0 for ONE environment using `self.np_random`."""
        raise NotImplementedError

    def next_vars(self, s_t):
        """Return [P_load.., P_gen_max.., aux..] for ONE environment given its state."""
        raise NotImplementedError

    # ---- batched versions (override for speed) ---------------------------------------------------
    def init_state_batch(self, indices):
        out = np.zeros((len(indices), self.state_N))
        for k, i in enumerate(indices):
            self._env_index = int(i)
            s0 = np.asarray(self.init_state(), dtype=np.float64)
            if s0.size != self.state_N:  # anm_env.py:273-277
                raise EnvInitializationError(
                    "Expected size of initial state s0 is %d but actual is %d" % (self.state_N, s0.size)
                )
            out[k] = s0
        self._env_index = 0
        return out

    def next_vars_batch(self, state):
        """TENSOR-LEVEL HOOK.  `state`: the [B, state_N] float64 CUDA tensor of all instances.  Return the
        [B, n_load + n_gen + K] stochastic variables of the next step as a CUDA tensor (torch ops, no host round trip) --
        or a NumPy array --, or None to use the built-in device table.  Override it in a custom environment that runs
        many instances: the default below calls the reference's per-environment hook `next_vars(s_t)` once per instance
        on the host (one device-to-host copy and B Python calls per step), which is what keeps an unmodified
        single-environment subclass (examples/simple_env.py) working."""
        if self.spec.table is not None:
            return None
        s = state.cpu().numpy()
        out = np.zeros((self.num_envs, self.spec.n_next_vars))
        for i in range(self.num_envs):
            self._env_index = i
            v = np.asarray(self.next_vars(s[i]), dtype=np.float64)
            if v.size != self.spec.n_next_vars:  # anm_env.py:371-374
                raise EnvNextVarsError(
                    "Next vars vector has size %d but expected is %d" % (v.size, self.spec.n_next_vars)
                )
            out[i] = v
        self._env_index = 0
        return out

    # ---- RNG (one PCG64 stream per env; global env g is seeded with seed + g, for any sharding) ----
    def _seed_rngs(self, seed):
        if seed is None:
            self._rngs = [_make_rng(None) for _ in range(self.num_envs)]
        else:
            self._rngs = [_make_rng(int(seed) + self.env_offset + i) for i in range(self.num_envs)]

    @property
    def np_random(self):
        if self._rngs is None:
            self._seed_rngs(None)
        return self._rngs[self._env_index]

    @property
    def unwrapped(self):
        return self

    @property
    def terminated(self):
        return self._term_u8.bool()

    # ---- reset (anm_env.py:235-311) ------------------------------------------------------------------
    def reset(self, *, seed=None, options=None, mask=None):
        """Reset all envs (or those selected by the boolean `mask`).  Returns (obs, {})."""
        if seed is not None:
            self._seed_rngs(seed)
        elif self._rngs is None:
            self._seed_rngs(None)
        B = self.num_envs
        todo = np.ones(B, dtype=bool) if mask is None else np.asarray(torch.as_tensor(mask).cpu(), dtype=bool).copy()
        if mask is None:
            self.timestep = 0
        s0 = np.zeros((B, self.state_N))
        conv = torch.zeros(B, dtype=torch.uint8, device=self.device)
        for attempt in range(N_INIT_STATES_MAX):
            idx = np.flatnonzero(todo)
            if idx.size == 0:
                break
            s0[idx] = self.init_state_batch(idx)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.native.reset(s0, mask=todo.astype(np.uint8), obs=self._obs, state=self.state, converged=conv)
            ok = conv.cpu().numpy().astype(bool)
            todo &= ~ok
        if todo.any():
            raise EnvInitializationError(
                "No non-terminal state found out of %d initial states for %d environment(s) of %s"
                % (N_INIT_STATES_MAX, int(todo.sum()), type(self).__name__)
            )
        sel = slice(None) if mask is None else torch.as_tensor(np.asarray(torch.as_tensor(mask).cpu(), dtype=bool), device=self.device)
        self._term_u8[sel] = 0
        self.e_loss[sel] = 0.0
        self.penalty[sel] = 0.0
        obs = self._observe()
        if self.observation_space is None:  # callable observation: bounds known only now (anm_env.py:297-300)
            n = obs.shape[-1]
            self.observation_space = Box(low=-np.ones(n) * np.inf, high=np.ones(n) * np.inf)
            self.observation_N = n
        return obs, {}

    def observation_batch(self, state):
        """TENSOR-LEVEL HOOK for callable observation spaces (anm_env.py:313-331 with a callable `observation`):
        return the [B, n_obs] observations computed from the [B, state_N] state tensor with torch ops, or None
        (default) to fall back to calling the per-environment callable row by row on the host."""
        return None

    def _observe(self):
        if self.spec.obs_callable is None:
            return self._obs
        ob = self.observation_batch(self.state)
        if ob is not None:
            return torch.as_tensor(ob, dtype=torch.float64, device=self.device)
        rows = [np.asarray(self.spec.obs_callable(s), dtype=np.float64) for s in self.state.cpu().numpy()]
        return torch.as_tensor(np.stack(rows), device=self.device)

    # ---- step (anm_env.py:333-453) -----------------------------------------------------------------------
    def check_action(self, action):
        ok = bool(((action >= self._act_lo) & (action <= self._act_hi)).all())
        assert ok, "Action %r invalid." % (action,)

    def step(self, action):
        """action: [num_envs, A] (MW / MVAr).  Returns (obs, reward, terminated, truncated, info)."""
        a = action if isinstance(action, torch.Tensor) else torch.as_tensor(np.asarray(action, dtype=np.float64))
        a = a.to(device=self.device, dtype=torch.float64).reshape(self.num_envs, -1).contiguous()
        if self.validate_actions:
            self.check_action(a)
        nv = self.next_vars_batch(self.state)
        self.native.step(a, nv, out=(self._obs, self.reward, self._term_u8), extras=self._extras)
        self.timestep += 1
        obs = self._observe()
        if self.spec.obs_callable is not None:  # terminal rows are zeros (anm_env.py:446-448)
            obs = obs * (~self.terminated).unsqueeze(1)
        truncated = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        return obs, self.reward, self.terminated, truncated, {}

    def observation(self, s_t=None):
        return self._observe()

    def render(self, mode="human"):
        raise NotImplementedError()

    def close(self):
        pass

    # ---- checkpoint / resume of the carried state ------------------------------
    def state_dict(self):
        """Everything a resumed run needs to continue bit-identically: the carried state of every instance (SoC, aux,
        terminated), the state / observation / cost tensors of the last step, the timestep, the host random streams
        (`np_random`, one PCG64 per instance) and -- for `device_init=True` environments -- the device-side streams."""
        soc, aux, term = self.native.get_state()
        d = {"soc": soc, "aux": aux, "terminated": term, "state": self.state.clone(), "obs": self._obs.clone(),
             "reward": self.reward.clone(), "e_loss": self.e_loss.clone(), "penalty": self.penalty.clone(),
             "timestep": self.timestep,
             "host_rng": None if self._rngs is None else [r.bit_generator.state for r in self._rngs]}  # fmt: skip
        if getattr(self, "_device_seeded", False):
            d["device_rng"] = self.native.get_rng()
        if self.track_full_state:
            d["full_state"] = self._full.clone()
        return d

    def load_state_dict(self, d):
        self.native.set_state(d["soc"], d["aux"], d["terminated"])
        self.state.copy_(d["state"])
        self._term_u8.copy_(d["terminated"])
        for name, buf in (("obs", self._obs), ("reward", self.reward), ("e_loss", self.e_loss), ("penalty", self.penalty)):
            if d.get(name) is not None:
                buf.copy_(d[name])
        self.timestep = int(d["timestep"])
        if d.get("host_rng") is not None:
            self._rngs = [_make_rng(0) for _ in d["host_rng"]]
            for r, st in zip(self._rngs, d["host_rng"]):
                r.bit_generator.state = st
        if d.get("device_rng") is not None:
            self.native.set_rng(d["device_rng"])
            self._device_seeded = True
        if self.track_full_state and d.get("full_state") is not None:
            self._full.copy_(d["full_state"])
