#!/usr/bin/env python
"""Establishes "the reference's reward level" for Env01 the way BASELINE.md §1 prescribes: PPO with SB3-default
hyper-parameters (n_steps 2048, batch 64, 10 epochs, lr 3e-4, ...; src/so100_mujoco_rl/main.py:56-64) on ONE CPU env,
here the fp64 oracle (MuJoCo / SB3 are not installable).  TEST INFRASTRUCTURE: it drives the plain-PyTorch `PPO`
learner on CPU tensors over the oracle; nothing of the product path runs.

    python tests/ref_level_ppo.py --seeds 0 1 2 --samples 1500000 --out profiles/r1_ref_level_env01.json
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_oracle  # noqa: E402
from so100_mujoco_rl_b200.ppo import PPO, PPOConfig  # noqa: E402


class OracleTorchEnv:
    """BatchedSo100Env's interface over the fp64 oracle, CPU tensors."""

    def __init__(self, task, n, seed):
        self.o = make_oracle(task, n, seed=seed)
        self.num_envs, self.obs_dim, self.act_dim, self.device = n, self.o.obs_dim, 6, torch.device("cpu")

    def reset(self):
        return torch.from_numpy(self.o.reset())

    def step(self, a):
        obs, rew, term, trunc, tobs, epr, epl = self.o.step(a.numpy(), nthreads=1)
        return SimpleNamespace(obs=torch.from_numpy(obs), reward=torch.from_numpy(rew.astype(np.float32)),
                               terminated=torch.from_numpy(term), truncated=torch.from_numpy(trunc), terminal_obs=torch.from_numpy(tobs),
                               ep_return=torch.from_numpy(epr.astype(np.float32)), ep_len=torch.from_numpy(epl))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="Env01")
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--samples", type=int, default=1_500_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_ref_level_env01.json"))
    args = ap.parse_args()
    torch.set_num_threads(2)
    from so100_mujoco_rl_b200.tasks import task_id
    runs = []
    for seed in args.seeds:
        env = OracleTorchEnv(task_id(args.env), 1, seed)
        algo = PPO(env, PPOConfig(n_steps=2048, n_minibatches=32, n_epochs=10, seed=seed, cuda_graph=False))  # batch 64 = 2048 / 32
        hist, t0 = [], time.time()
        algo.learn(args.samples, log_every=0, callback=hist.append)
        curve = [{"samples": h["samples"], "mean_step_reward": h["mean_step_reward"], "ep_return_mean": h["ep_return_mean"]} for h in hist]
        tail = [h["mean_step_reward"] for h in hist[-max(1, len(hist) // 10):]]
        runs.append({"seed": seed, "wall_s": time.time() - t0, "plateau_mean_step_reward": float(np.mean(tail)),
                     "plateau_episode_return": float(np.mean(tail)) * 4000, "curve": curve[:: max(1, len(curve) // 150)]})
        print(json.dumps({k: runs[-1][k] for k in ("seed", "wall_s", "plateau_mean_step_reward", "plateau_episode_return")}), flush=True)
    plate = [r["plateau_mean_step_reward"] for r in runs]
    out = {"env": args.env, "setup": "PPO, SB3-default hyper-parameters, 1 env, fp64 oracle, plain PyTorch learner on CPU",
           "samples_per_seed": args.samples, "plateau_mean_step_reward_mean": float(np.mean(plate)),
           "plateau_mean_step_reward_min_max": [float(min(plate)), float(max(plate))], "runs": runs}
    json.dump(out, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
