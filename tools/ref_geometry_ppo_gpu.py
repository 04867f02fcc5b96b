#!/usr/bin/env python
"""The counterpart of tests/ref_level_ppo.py on the product path: the SAME learner geometry the reference uses (SB3
defaults: ONE env, n_steps 2048, batch 64, 10 epochs) but on the CUDA env with the fused learner, 3 seeds, so that the
plateaus can be compared like for like.  python tools/ref_geometry_ppo_gpu.py --samples 1000000"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="Env01")
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--samples", type=int, default=1_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_geometry_gpu_env01.json"))
    args = ap.parse_args()
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    from so100_mujoco_rl_b200.ppo import FusedPPO, PPOConfig
    runs = []
    for seed in args.seeds:
        env = BatchedSo100Env(args.env, 1, device=0, seed=seed)
        algo = FusedPPO(env, PPOConfig(n_steps=2048, n_minibatches=32, n_epochs=10, seed=seed))
        hist, t0 = [], time.time()
        algo.learn(args.samples, log_every=0, callback=hist.append)
        tail = [h["mean_step_reward"] for h in hist[-max(1, len(hist) // 10):]]
        runs.append({"seed": seed, "wall_s": time.time() - t0, "plateau_mean_step_reward": float(np.mean(tail)),
                     "best_iteration_mean_step_reward": max(h["mean_step_reward"] for h in hist),
                     "curve": [{"samples": h["samples"], "mean_step_reward": h["mean_step_reward"]} for h in hist[:: max(1, len(hist) // 150)]]})
        print(json.dumps({k: runs[-1][k] for k in ("seed", "wall_s", "plateau_mean_step_reward", "best_iteration_mean_step_reward")}), flush=True)
        env.close()
    plate = [r["plateau_mean_step_reward"] for r in runs]
    out = {"env": args.env, "setup": "FusedPPO, SB3-default geometry (1 env, n_steps 2048, batch 64, 10 epochs), CUDA env",
           "samples_per_seed": args.samples, "plateau_mean_step_reward_mean": float(np.mean(plate)),
           "plateau_mean_step_reward_min_max": [float(min(plate)), float(max(plate))], "runs": runs}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
