import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def spec():
    from so100_mujoco_rl_b200.model import load_model
    return load_model()


@pytest.fixture(scope="session")
def native_lib():
    """libso100_b200.so, built in-tree if nvcc is present and the sources are newer."""
    from so100_mujoco_rl_b200 import _build, _native
    try:
        _build.build_native()
    except Exception:
        if not os.path.exists(_native.LIB_PATH):
            raise
    return _native.lib()


def make_oracle(task, n, seed=0, env_offset=0, flags=0, max_episode_steps=None, spec=None):
    from oracle.pyoracle import Oracle
    from so100_mujoco_rl_b200.model import load_model
    from so100_mujoco_rl_b200.tasks import make_task_cfg
    spec = spec or load_model()
    return Oracle(spec.to_ctypes(), make_task_cfg(task, n, seed=seed, env_offset=env_offset, flags=flags,
                                                  max_episode_steps=max_episode_steps))
