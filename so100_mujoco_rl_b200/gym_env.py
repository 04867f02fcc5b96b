"""`So100Env` — the single-environment Gymnasium face (`reset(seed=...) -> (obs, info)`, `step(a) -> (obs, reward,
terminated, truncated, info)`) over the batched simulator with one env, for tooling written against the reference's
`gym.make("Env01")` objects (src/so100_mujoco_rl/main.py:85, :131, :182; registration in __init__.py:5-45).

The simulator resets a finished env inside the same kernel launch (VecEnv semantics); this wrapper turns that back
into Gymnasium's: the step that ends an episode returns the episode's LAST observation, and the following `reset()`
hands out the first observation of the next episode, which the kernel has already produced.
If `gymnasium` is importable the ids Env01-v1 / Env02-v1 / Env05-v1 / Env06-v1 can be registered with `register()`.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .vec_env import So100VecEnv


class So100Env:
    metadata = So100VecEnv.metadata

    def __init__(self, env_id: str | int, device: int = 0, seed: int = 0, render_mode: str | None = None, backend: Any = None,
                 max_episode_steps: int | None = None):
        self._vec = So100VecEnv(env_id, 1, device=device, seed=seed, backend=backend, max_episode_steps=max_episode_steps,
                                clip_actions=False)  # the reference env does not clip (SB3 does, before calling it)
        self.observation_space, self.action_space = self._vec.observation_space, self._vec.action_space
        self.render_mode = render_mode
        self._pending: np.ndarray | None = None  # first obs of the next episode, produced by the in-kernel auto-reset
        self._needs_reset = True

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self._vec.seed(seed)
            self._pending = None
        obs = self._pending if self._pending is not None else self._vec.reset()[0]
        self._pending, self._needs_reset = None, False
        self._last = np.array(obs, dtype=np.float32)
        return self._last.copy(), {}

    def step(self, action):
        if self._needs_reset:
            raise RuntimeError("step() called before reset() (gymnasium OrderEnforcing)")
        obs, rew, dones, infos = self._vec.step(np.asarray(action, dtype=np.float32).reshape(1, 6))
        info = dict(infos[0])
        terminated = truncated = False
        out = obs[0]
        if dones[0]:
            truncated = bool(info.pop("TimeLimit.truncated", False))
            terminated = not truncated
            self._pending, self._needs_reset = obs[0].copy(), True
            out = info.pop("terminal_observation")
        self._last = np.array(out, dtype=np.float32)
        return self._last.copy(), float(rew[0]), terminated, truncated, info

    # ---- the reference's getters (envs/env_base_01.py:107-142, env_base_02.py:85-86), served from the last observation:
    #      they read the same mjData fields the observation is assembled from
    def get_joint_angles(self) -> np.ndarray:
        return self._last[:6].copy()  # Env05: the COMMANDED angles, as env_base_02.py:85-86 returns

    def _need15(self):
        if self._last.shape[0] != 15:
            raise AttributeError("Env05 observes the projected cube centre, not Cartesian positions (env05_v1.py:32-75)")

    def get_block_pos(self) -> np.ndarray:
        self._need15()
        return self._last[9:12].copy()

    def get_end_effector_pos(self) -> np.ndarray:
        self._need15()
        return self._last[12:15].copy()

    def get_block_to_end_distance(self) -> float:
        self._need15()
        return float(np.linalg.norm(self._last[6:9]))

    def render(self):
        return None  # rendering is outside the hot path

    def close(self):
        self._vec.close()

    @property
    def unwrapped(self):
        return self


def register() -> list[str]:
    """Register EnvNN-v1 ids with gymnasium (if installed) pointing at So100Env; returns the ids registered."""
    try:
        import gymnasium as gym  # type: ignore
    except Exception:  # noqa: BLE001
        return []
    ids = []
    for name in ("Env01", "Env02", "Env05", "Env06"):
        gid = f"{name}-b200-v1"
        gym.register(id=gid, entry_point="so100_mujoco_rl_b200.gym_env:So100Env", kwargs={"env_id": name}, disable_env_checker=True)
        ids.append(gid)
    return ids
