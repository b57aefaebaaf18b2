"""Minimal stand-in for `gymnasium` (TEST INFRASTRUCTURE ONLY, see oracle/README.md).

Only what the reference touches: `Env` seeding (reference anm_env.py:116,257),
`spaces.Box` (:231,299,493), `envs.registration.register` (gym_anm/__init__.py:8).
Seeding follows gymnasium 1.0 `utils.seeding.np_random`:
Generator(PCG64(SeedSequence(seed))).
"""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration as _registration

__version__ = "0.0-standin"


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


def make(env_id, **kwargs):
    if ":" in env_id:
        module, env_id = env_id.split(":")
        importlib.import_module(module)
    entry = _registration.registry[env_id]
    mod_name, cls_name = entry.split(":")
    return getattr(importlib.import_module(mod_name), cls_name)(**kwargs)
