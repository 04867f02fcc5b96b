"""Observation / action spaces of Env01, Env02, Env05 with the reference's bounds.

Reference: So100BaseEnv.get_observation_space (envs/env_base_01.py:63-75), _set_action_space (:77-83),
So100OffscreenBaseEnv.get_observation_space (envs/env_base_02.py:56-69).  When gymnasium is importable the real
`gymnasium.spaces.Box` is returned (what SB3 expects); otherwise a structural stand-in with the same attributes.
"""
from __future__ import annotations

import numpy as np

from .model import ModelSpec
from .tasks import TASK_ENV05


class _Box:
    """Minimal stand-in for gymnasium.spaces.Box (low/high/shape/dtype/sample/contains)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def _box(low, high):
    try:
        from gymnasium.spaces import Box  # type: ignore
        return Box(np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32), dtype=np.float32)
    except Exception:
        return _Box(low, high)


def observation_space(task: int, spec: ModelSpec):
    mins, maxs = list(spec.jnt_range[:, 0]), list(spec.jnt_range[:, 1])
    if task == TASK_ENV05:
        return _box([*mins, 0.0, 0.0], [*maxs, 5.0, 5.0])
    return _box([*mins, -1.0, -1.0, -1.0, *([-0.5] * 6)], [*maxs, 1.0, 1.0, 1.0, *([0.5] * 6)])


def action_space():
    return _box([-1.0] * 6, [1.0] * 6)
