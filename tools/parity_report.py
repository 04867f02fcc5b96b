#!/usr/bin/env python
"""GPU-vs-oracle parity report: N envs x K steps per task, same seeds and actions; writes gpurun_out/parity.json.

    python tools/parity_report.py [--envs 256] [--steps 1000]

Reports max / quantiles of |d qpos|, |d qvel| (every step), |d obs| per column and |d reward|.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    args = ap.parse_args()
    import torch
    from oracle.pyoracle import Oracle
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    from so100_mujoco_rl_b200.model import load_model
    from so100_mujoco_rl_b200.tasks import make_task_cfg

    spec = load_model()
    report = {}
    for task in (1, 2, 5, 6):
        n = args.envs
        env = BatchedSo100Env(task, n, device=0, seed=11)
        o = Oracle(spec.to_ctypes(), make_task_cfg(task, n, seed=11))
        env.reset(); o.reset(nthreads=0)
        rng = np.random.default_rng(5)
        dq_all, dv_all, dr_all, dobs_cols = [], [], [], []
        flag_mismatch = 0
        alive = np.ones(n, dtype=bool)   # False once an env's done flags differed: from then on the two runs are in different episodes
        for t in range(args.steps):
            a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
            r = env.step(torch.from_numpy(a).cuda())
            oo, ro, to, co, *_ = o.step(a, nthreads=0)
            st = env.get_state()
            mism = (r.terminated.cpu().numpy() != to) | (r.truncated.cpu().numpy() != co)
            flag_mismatch += int((mism & alive).sum())
            alive &= ~mism
            w = np.where(alive, 1.0, 0.0)   # statistics over the envs still in the same episode in both runs
            dq_all.append(np.abs(st["qpos"].cpu().numpy().T - o.gather("qpos")).max(axis=1) * w)
            dv_all.append(np.abs(st["qvel"].cpu().numpy().T - o.gather("qvel")).max(axis=1) * w)
            dr_all.append(np.abs(r.reward.cpu().numpy() - ro) * w)
            dobs_cols.append((np.abs(r.obs.cpu().numpy() - oo) * w[:, None]).max(axis=0))
        dq, dv, dr = np.array(dq_all), np.array(dv_all), np.array(dr_all)
        q = lambda x, p: float(np.quantile(x, p))  # noqa: E731
        report[f"Env0{task}"] = {
            "envs": n, "steps": args.steps,
            "dqpos": {"max": float(dq.max()), "p999": q(dq, .999), "p99": q(dq, .99), "median": q(dq, .5), "final_max": float(dq[-1].max())},
            "dqvel": {"max": float(dv.max()), "p999": q(dv, .999), "p99": q(dv, .99), "median": q(dv, .5), "final_max": float(dv[-1].max())},
            "dreward": {"max": float(dr.max()), "p999": q(dr, .999), "p99": q(dr, .99), "median": q(dr, .5)},
            "dobs_col_max": [float(x) for x in np.array(dobs_cols).max(axis=0)],
            "worst_step_dq": int(dq.max(axis=1).argmax()), "worst_env_dq": int(dq.max(axis=0).argmax()),
            "done_flag_mismatches": flag_mismatch, "envs_bifurcated": int((~alive).sum()),
            "samples_dq_over_2e-5": int((dq > 2e-5).sum()), "envs_ever_dq_over_2e-5": int((dq > 2e-5).any(axis=0).sum()),
            "samples": int(dq.size),
            "stats": env.stats(),
        }
        print(task, json.dumps(report[f"Env0{task}"]), flush=True)
        env.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
