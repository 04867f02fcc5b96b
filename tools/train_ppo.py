#!/usr/bin/env python
"""Trains the GPU PPO on a batched so100 env and writes the learning curve (gpurun_out/ppo_<env>.json).

    python tools/train_ppo.py --env Env01 --num-envs 4096 --samples 30000000
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="Env01")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--samples", type=float, default=3e7)
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-episode-steps", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    import torch
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    from so100_mujoco_rl_b200.ppo import PPO, PPOConfig
    env = BatchedSo100Env(args.env, args.num_envs, device=0, seed=args.seed, max_episode_steps=args.max_episode_steps or None)
    algo = PPO(env, PPOConfig(n_steps=args.n_steps, n_minibatches=args.minibatches, n_epochs=args.epochs, seed=args.seed))
    hist = []

    def cb(rec):
        hist.append(rec)
        if rec["iter"] % 20 == 0 or rec["iter"] == 1:
            print(json.dumps(rec), flush=True)
    t0 = time.time()
    st = algo.learn(int(args.samples), log_every=0, callback=cb)
    wall = time.time() - t0
    out = {"env": args.env, "num_envs": args.num_envs, "n_steps": args.n_steps, "samples": st.samples, "wall_s": wall,
           "rollout_s": st.rollout_s, "update_s": st.update_s, "samples_per_s": st.samples / wall,
           "rollout_env_steps_per_s": st.samples / st.rollout_s, "history": hist[:: max(1, len(hist) // 200)] + hist[-1:],
           "kernel_variant": env.kernel_variant, "stats": env.stats()}
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, f"ppo_{args.env}_s{args.seed}.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: out[k] for k in ("samples", "wall_s", "rollout_s", "update_s", "samples_per_s")})


if __name__ == "__main__":
    main()
