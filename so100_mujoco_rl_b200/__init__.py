"""so100_mujoco_rl_b200 — B200-native batched simulator for the so100 arm tasks (Env01 / Env02 / Env05).

Only the hot path of PieterBecking/so100-mujoco-rl lives here: batched reset/step behind the SB3 VecEnv seam.
Importing the package does not need a GPU; constructing an env does (there is no CPU fallback).
"""
from .model import ModelSpec, load_model  # noqa: F401
from .tasks import TASK_ENV01, TASK_ENV02, TASK_ENV05, make_task_cfg, task_id  # noqa: F401
from .vec_env import So100VecEnv  # noqa: F401

__all__ = ["ModelSpec", "load_model", "make_task_cfg", "task_id", "So100VecEnv", "So100Env", "BatchedSo100Env",
           "TASK_ENV01", "TASK_ENV02", "TASK_ENV05"]


def __getattr__(name):  # BatchedSo100Env imports torch; keep `import so100_mujoco_rl_b200` light
    if name == "BatchedSo100Env":
        from .batched_env import BatchedSo100Env
        return BatchedSo100Env
    if name == "So100Env":
        from .gym_env import So100Env
        return So100Env
    raise AttributeError(name)
