"""State / observation vector specifications -> flat `anm_var_spec` arrays.

Host-side half of ANMEnv.__init__ (reference gym_anm/envs/anm_env.py:139-156, 193-233,
475-549): expanding 'all', default units, observation bounds.  Unit conversions follow
Simulator._gather_state (simulator.py:551-636).
"""
import numpy as np

from ._capi import VAR_SPEC_DTYPE
from .errors import ObsNotSupportedError, ObsSpaceError, UnitsNotSupportedError
from .network_spec import STATE_VARIABLES

QUANTITY = {k: i for i, k in enumerate(
    ["bus_p", "bus_q", "bus_v_magn", "bus_v_ang", "bus_i_magn", "bus_i_ang", "dev_p", "dev_q", "des_soc",
     "gen_p_max", "branch_p", "branch_q", "branch_s", "branch_i_magn", "branch_i_ang", "aux"])}  # fmt: skip


def expand_all_ids(cn, values, K):
    """'all' -> explicit id lists (anm_env.py:523-549)."""
    if values is None:
        return None
    out = []
    for o in values:
        key, ids = o[0], o[1]
        unit = o[2] if len(o) > 2 else None
        if isinstance(ids, str) and ids == "all":
            if "bus" in key:
                ids = list(cn.buses.keys())
            elif "dev" in key:
                ids = list(cn.devices.keys())
            elif "des" in key:
                ids = list(cn.des_ids)
            elif "gen" in key:
                ids = list(cn.gen_ids)
            elif "branch" in key:
                ids = list(cn.branches.keys())
            elif key == "aux":
                ids = list(range(K))
            else:
                raise ObsNotSupportedError(key, STATE_VARIABLES.keys())
        out.append((key, ids, unit))
    return out


def default_units(values):
    """Fill in the default unit of 2-tuples (anm_env.py:505-510)."""
    out = []
    for o in values:
        o = tuple(o)
        if len(o) == 2:
            o = o + (STATE_VARIABLES[o[0]][0],)
        out.append(o)
    return out


def check_observation_vars(observation, state_bounds, K):
    """Validation of a list-style observation spec (envs/utils.py:71-117)."""
    for obs in observation:
        if len(obs) not in (2, 3):
            raise ObsSpaceError("The observation tuple {} should have 2 or 3 elements.".format(obs))
        key, nodes = obs[0], obs[1]
        if key not in STATE_VARIABLES:
            raise ObsNotSupportedError(key, STATE_VARIABLES)
        if isinstance(nodes, str) and nodes == "all":
            pass
        elif key == "aux":
            for n in nodes:
                if n >= K:
                    raise ObsSpaceError("Aux variable index {} is out of bound for {} aux variables.".format(n, K))
        elif isinstance(nodes, list):
            for n in nodes:
                if n not in state_bounds[key]:
                    raise ObsSpaceError("Observation {} is not supported for device/branch/bus with ID {}.".format(key, n))
        else:
            raise ObsSpaceError()
        if len(obs) == 3 and obs[2] not in STATE_VARIABLES[key]:
            raise UnitsNotSupportedError(obs[2], STATE_VARIABLES[key], key)


def _unit_factors(cn, key, ident, unit):
    m = cn.baseMVA
    if unit in (None, "pu", "rad"):
        return 1.0, 1.0
    if unit in ("MW", "MVAr", "MVA", "MWh"):
        return m, 1.0
    if unit == "degree":
        return 180.0, np.pi
    if unit == "kV":
        return cn.buses[ident].baseKV, 1.0
    if unit == "kA":
        return m, cn.buses[ident].baseKV
    raise UnitsNotSupportedError(unit, STATE_VARIABLES[key], key)


def build_var_specs(cn, values, K, aux_bounds=None, with_bounds=False):
    """(key, ids, unit) triples -> array of VAR_SPEC_DTYPE (+ low / high when `with_bounds`)."""
    bus_pos = {b: k for k, b in enumerate(cn.buses)}
    dev_pos = {d: k for k, d in enumerate(cn.devices)}
    br_pos = {b: k for k, b in enumerate(cn.branches)}
    rows = []
    for key, ids, unit in values:
        if key not in QUANTITY:
            raise ObsNotSupportedError(key, STATE_VARIABLES.keys())
        for ident in ids:
            if key == "aux":
                pos, mul, div = int(ident), 1.0, 1.0
                lo, hi = (-np.inf, np.inf) if aux_bounds is None else (aux_bounds[ident][0], aux_bounds[ident][1])
            else:
                if "bus" in key:
                    pos = bus_pos[ident]
                elif "branch" in key:
                    pos = br_pos[tuple(ident)]
                else:
                    pos = dev_pos[ident]
                mul, div = _unit_factors(cn, key, ident, unit)
                lo, hi = cn.state_bounds[key][ident if "branch" not in key else tuple(ident)][unit]
            if not with_bounds:
                lo, hi = -np.inf, np.inf
            rows.append((QUANTITY[key], pos, mul, div, lo, hi))
    return np.array(rows, dtype=VAR_SPEC_DTYPE)


def check_env_args(K, delta_t, lamb, gamma, observation, aux_bounds, state_bounds):
    """Constructor-argument validation (reference envs/utils.py:7-68)."""
    from .errors import ArgsError

    if K < 0:
        raise ArgsError("The argument K is %d but should be >= 0." % K)
    if delta_t <= 0:
        raise ArgsError("The argument delta_t is %.2f but should be > 0." % delta_t)
    if lamb < 0:
        raise ArgsError("The argument lamb is %d but should be >= 0." % lamb)
    if gamma < 0 or gamma > 1:
        raise ArgsError("The argument gamma is %.4f but should be in [0, 1]." % gamma)
    if isinstance(observation, str) and observation == "state":
        pass
    elif isinstance(observation, list):
        check_observation_vars(observation, state_bounds, K)
    elif callable(observation):
        pass
    else:
        raise ArgsError(
            'The argument observation is of type {} but should be a list, a callable, or "state".'.format(
                type(observation)
            )
        )
    if aux_bounds is not None and len(aux_bounds) != K:
        raise ArgsError("The argument aux_bounds has length {} but K={}.".format(len(aux_bounds), K))


class HostEnvSpec:
    """Everything ANMEnv.__init__ derives from its arguments, in flat form.

    network, observation, K, delta_t, gamma, lamb, aux_bounds, costs_clipping have the
    meaning of reference anm_env.py:79-113.  `table` ([T, n_load + n_gen], MW) optionally
    enables the built-in periodic next_vars (ANM6Easy)."""

    def __init__(self, network, observation, K, delta_t, gamma, lamb, aux_bounds=None, costs_clipping=None,
                 table=None):  # fmt: skip
        from .network_spec import CompiledNetwork

        self.K, self.gamma, self.lamb, self.delta_t, self.aux_bounds = K, gamma, lamb, delta_t, aux_bounds
        if costs_clipping is None:
            c1, c2 = np.inf, np.inf
        else:
            c1 = np.inf if costs_clipping[0] is None else costs_clipping[0]
            c2 = np.inf if costs_clipping[1] is None else costs_clipping[1]
        self.costs_clipping = (c1, c2)
        self.cn = CompiledNetwork(network, delta_t, lamb)
        check_env_args(K, delta_t, lamb, gamma, observation, aux_bounds, self.cn.state_bounds)
        cn = self.cn
        self.state_values = expand_all_ids(
            cn,
            [("dev_p", "all", "MW"), ("dev_q", "all", "MVAr"), ("des_soc", "all", "MWh"), ("gen_p_max", "all", "MW"),
             ("aux", "all", None)],
            K,
        )  # fmt: skip
        self.state_N = sum(len(s[1]) for s in self.state_values)
        self.state_specs = build_var_specs(cn, self.state_values, K)
        # observation space (anm_env.py:497-521)
        self.obs_callable = None
        if isinstance(observation, str) and observation == "state":
            self.obs_values = [tuple(v) for v in self.state_values]
        elif isinstance(observation, list):
            self.obs_values = expand_all_ids(cn, default_units(observation), K)
        elif callable(observation):
            self.obs_values, self.obs_callable = None, observation
        else:
            raise ObsSpaceError()
        if self.obs_values is not None:
            self.obs_specs = build_var_specs(cn, self.obs_values, K, aux_bounds, with_bounds=True)
            self.obs_low, self.obs_high = self.obs_specs["low"].copy(), self.obs_specs["high"].copy()
        else:  # callable: the device emits the (unclipped) state vector, the callable runs on the host
            self.obs_specs, self.obs_low, self.obs_high = self.state_specs, None, None
        # action space (anm_env.py:475-495, simulator.py:341-380)
        m = cn.baseMVA
        lo, hi = [], []
        for ids, a, b in ((cn.gen_ids, "p_min", "p_max"), (cn.gen_ids, "q_min", "q_max"),
                          (cn.des_ids, "p_min", "p_max"), (cn.des_ids, "q_min", "q_max")):  # fmt: skip
            for i in ids:
                lo.append(getattr(cn.devices[i], a) * m)
                hi.append(getattr(cn.devices[i], b) * m)
        self.action_low, self.action_high = np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)
        self.table = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
        self.n_next_vars = cn.N_load + cn.N_non_slack_gen + K
        self.n_full_state = 6 * cn.N_bus + 4 * cn.N_device + 5 * cn.N_branch + K

    def descs(self):
        from ._capi import make_descs

        return make_descs(self.cn.flat(), self.cn.baseMVA, self.delta_t, self.lamb, self.K, self.gamma,
                          self.costs_clipping, self.state_specs, self.obs_specs, self.table)  # fmt: skip

    def full_state_slices(self):
        """name -> slice into a full-state row (order of include/anm_b200.h)."""
        N, D, L, K = self.cn.N_bus, self.cn.N_device, self.cn.N_branch, self.K
        out, off = {}, 0
        for name, n in (("bus_p", N), ("bus_q", N), ("bus_v_magn", N), ("bus_v_ang", N), ("bus_i_magn", N),
                        ("bus_i_ang", N), ("dev_p", D), ("dev_q", D), ("des_soc", D), ("gen_p_max", D),
                        ("branch_p", L), ("branch_q", L), ("branch_s", L), ("branch_i_magn", L),
                        ("branch_i_ang", L), ("aux", K)):  # fmt: skip
            out[name] = slice(off, off + n)
            off += n
        return out


def anm6easy_spec():
    """The ANM6Easy-v0 configuration (reference anm6_easy.py:11-23)."""
    from .networks import anm6_network, anm6easy_tables

    loads, gens = anm6easy_tables()
    table = np.ascontiguousarray(np.vstack((loads, gens)).T)  # [96, 5]: P1 P3 P5 | P2 P4
    delta_t = 0.25
    return HostEnvSpec(anm6_network(), "state", 1, delta_t, 0.995, 100, np.array([[0, 24 / delta_t - 1]]), (1, 100),
                       table=table)  # fmt: skip
