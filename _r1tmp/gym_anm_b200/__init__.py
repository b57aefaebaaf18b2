"""B200-native batched gym-anm step() path.

Public surface (mirrors reference gym_anm/__init__.py and gym_anm/envs/__init__.py):
`ANM6Easy` (single env, reference call shapes), `BatchedANM6Easy`, `BatchedANMEnv`,
the network dict schema and exception types.  `ANM6Easy-v0` is registered with
Gymnasium when Gymnasium is installed.
"""
from .errors import *  # noqa: F401,F403
from .network_spec import CompiledNetwork  # noqa: F401
from .env_spec import HostEnvSpec, anm6easy_spec  # noqa: F401

__all__ = ["ANM6Easy", "BatchedANM6Easy", "BatchedANMEnv", "BatchedSimulator", "CompiledNetwork", "HostEnvSpec",
           "MPCAgentConstant", "MPCAgentPerfect"]


def __getattr__(name):  # torch-dependent modules are imported lazily
    if name in ("ANM6Easy", "BatchedANM6Easy"):
        from . import anm6

        return getattr(anm6, name)
    if name == "BatchedANMEnv":
        from .anm_env import BatchedANMEnv

        return BatchedANMEnv
    if name in ("MPCAgentConstant", "MPCAgentPerfect"):
        from . import agents

        return getattr(agents, name)
    if name == "BatchedSimulator":
        from .simulator import BatchedSimulator

        return BatchedSimulator
    raise AttributeError(name)


try:  # reference gym_anm/__init__.py:8-11
    from gymnasium.envs.registration import register

    register(id="ANM6Easy-v0", entry_point="gym_anm_b200.anm6:ANM6Easy")
except Exception:  # noqa: BLE001  (gymnasium not installed in this image)
    pass
