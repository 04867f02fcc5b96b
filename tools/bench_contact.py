#!/usr/bin/env python
"""Step time of the arm-floor contact physics (SO100_FLAG_ARM_CONTACT) beside the default, same workload as bench.py:
Env01, U(-1,1) actions, decorrelated start.  One JSON line per (flags, envs)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from so100_mujoco_rl_b200.batched_env import BatchedSo100Env  # noqa: E402


def run(task, n, flags, steps, warm):
    env = BatchedSo100Env(task, n, device=0, seed=1, flags=flags)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = (torch.rand(8, n, 6, device="cuda", generator=g) * 2 - 1).contiguous()
    for i in range(warm):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        env.step(acts[i % 8])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    st = env.stats() if hasattr(env, "stats") else None
    env.close()
    return {"task": task, "envs": n, "flags": flags, "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "ms_max": ms[-1],
            "env_steps_per_s": n / (ms[len(ms) // 2] * 1e-3), "stats": st}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[65536])
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--flags", type=int, nargs="+", default=[0, 16])
    ap.add_argument("--task", type=int, default=1)
    a = ap.parse_args()
    for n in a.envs:
        for f in a.flags:
            print(json.dumps(run(a.task, n, f, a.steps, a.warmup)), flush=True)
