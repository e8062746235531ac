"""gym surface objects of the environment, without depending on ``gym``.

The reference declares ``action_space = gym.spaces.Box(low=-pi, high=pi, shape=(1,))`` and
``observation_space = gym.spaces.Box(low=-neighborDist, high=neighborDist, shape=(laser_num*4,))``
(collision_avoidance/envs/collision_avoidence_env.py:52-53) and registers the env under the id
``collision_avoidance-v0`` (collision_avoidance/__init__.py:1-6).  ``Box`` below offers the part
of gym's Box that RL libraries touch (``low/high/shape/dtype/sample/contains``); ``register``
also registers with a real ``gym`` / ``gymnasium`` when one is importable.
"""
from __future__ import annotations

from math import pi
from typing import Callable, Dict, Optional, Tuple

import numpy as np

ENV_ID = "collision_avoidance-v0"                      # collision_avoidance/__init__.py:4
ENTRY_POINT = "collision_avoidance_b200.envs:Collision_Avoidance_Env"
registry: Dict[str, str] = {}


class Box:
    """Axis-aligned box in R^n with scalar bounds, as the reference constructs it."""

    def __init__(self, low: float, high: float, shape: Tuple[int, ...], dtype=np.float32, seed: Optional[int] = None):
        if not low <= high:
            raise ValueError("Box needs low <= high")
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self, batch_shape: Tuple[int, ...] = ()) -> np.ndarray:
        return self._rng.uniform(self.low, self.high, size=tuple(batch_shape) + self.shape).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape[-len(self.shape):] == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __eq__(self, other):
        return isinstance(other, Box) and (self.low, self.high, self.shape) == (other.low, other.high, other.shape)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def env_spaces(neighbor_dist: float = 1.5, laser_num: int = 16) -> Tuple[Box, Box]:
    """(action_space, observation_space) of collision_avoidence_env.py:52-53."""
    return Box(-pi, pi, (1,)), Box(-neighbor_dist, neighbor_dist, (laser_num * 4,))


def register(id: str = ENV_ID, entry_point: str = ENTRY_POINT) -> None:
    """Record the id here and, when a gym package exists, in its registry as well."""
    registry[id] = entry_point
    for mod in ("gymnasium", "gym"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"])
            reg.register(id=id, entry_point=entry_point)
        except Exception:       # absent package or id already registered: our own registry still holds it
            pass


def make(id: str = ENV_ID, **kwargs):
    """``gym.make``-style construction from the local registry."""
    if id not in registry:
        raise KeyError(f"no environment registered under {id!r}")
    module, _, attr = registry[id].partition(":")
    ctor: Callable = getattr(__import__(module, fromlist=[attr]), attr)
    return ctor(**kwargs)


register()
