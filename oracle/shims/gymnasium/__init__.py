"""Minimal stand-in for the `gymnasium` names the reference touches.

TEST INFRASTRUCTURE ONLY. gymnasium is not installed in this image and there is no network, so the
live reference at /root/reference can only be imported (to generate golden vectors, see
oracle/gen_golden.py) through this shim. Nothing in the product package imports it.
"""
import numpy as np

from . import spaces, utils, envs  # noqa: F401
from .envs.registration import register, make, registry  # noqa: F401


class Env:
    metadata = {}
    render_mode = None
    action_space = None
    observation_space = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        # gymnasium.Env.reset reseeds self._np_random only when a seed is given
        if seed is not None:
            self._np_random, _ = utils.seeding.np_random(seed)
        return None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = utils.seeding.np_random(None)
        return self._np_random

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)
