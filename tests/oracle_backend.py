"""CPU stand-in for the adapter's backend protocol (reset_np / step_np / close), driven by the fp64 oracle.
TEST-ONLY: lets the VecEnv host logic be exercised without a GPU.  The product backend is TorchBackend."""
import numpy as np

from conftest import make_oracle


class OracleBackend:
    def __init__(self, task, n, seed=0, max_episode_steps=None, env_offset=0):
        self.o = make_oracle(task, n, seed=seed, max_episode_steps=max_episode_steps, env_offset=env_offset)

    def reset_np(self):
        return self.o.reset()

    def step_np(self, actions):
        obs, rew, term, trunc, tobs, epr, epl = self.o.step(actions)
        return obs, rew.astype(np.float32), term, trunc, tobs, epr.astype(np.float32), epl

    def close(self):
        self.o.close()
