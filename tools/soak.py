#!/usr/bin/env python
"""Soak run: N envs x many steps of U(-1,1) actions per task; reports the library's health counters (non-finite resets,
substeps whose last solver sweep still moved qacc by > 2e-3) and state bounds.  python tools/soak.py --steps 100000"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=100000)
    ap.add_argument("--flags", type=int, default=0, help="SO100_FLAG_* bits (16 = arm-floor contact)")
    ap.add_argument("--tasks", type=int, nargs="+", default=[1, 2, 5, 6])
    args = ap.parse_args()
    import torch
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    ring = [torch.rand((args.envs, 6), device=dev, generator=g) * 2 - 1 for _ in range(97)]
    out = {}
    for task in args.tasks:
        env = BatchedSo100Env(task, args.envs, device=0, seed=5, flags=args.flags)
        env.reset()
        dones = torch.zeros((), device=dev, dtype=torch.int64)
        rsum = torch.zeros((), device=dev, dtype=torch.float64)
        for t in range(args.steps):
            r = env.step(ring[t % 97])
            if t % 64 == 0:
                dones += (r.terminated | r.truncated).sum()
                rsum += r.reward.double().sum()
        st, s = env.get_state(), env.stats()
        out[f"Env0{task}"] = {"env_steps": args.envs * args.steps, "flags": args.flags, **s, "qpos_abs_max": float(st["qpos"].abs().max()),
                              "qvel_abs_max": float(st["qvel"].abs().max()), "all_finite": bool(torch.isfinite(st["qpos"]).all() and torch.isfinite(st["qvel"]).all()),
                              "sampled_mean_reward": float(rsum) / (args.envs * ((args.steps + 63) // 64)), "sampled_dones": int(dones)}
        env.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
