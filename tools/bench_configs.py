#!/usr/bin/env python
"""BASELINE configs 2-4 at 65 536 envs on one GPU, with the events each config is meant to exercise (SURVEY.md §8 d):

  config 2  Env01, random actions                                   -> env-steps/s
  config 3  Env02, random actions + a scripted subset servoed onto the block (damped Jacobian-transpose steps from the
            library's own kinematics entry point) so that reach -> bonus -> relocate fires mid-episode -> relocations/s
  config 4  Env05, random actions: the lost-cube counter terminates episodes continuously -> resets/s

Step time is measured with CUDA events around `env.step` only (the scripted controller is host-driven tooling).
    python tools/bench_configs.py [--envs 65536] [--steps 300] [--scripted 2048]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--scripted", type=int, default=2048)
    args = ap.parse_args()
    import torch
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    n, out = args.envs, {}
    for cfg, task in (("config2_env01", 1), ("config3_env02", 2), ("config4_env05", 5)):
        env = BatchedSo100Env(task, n, device=0, seed=0)
        obs = env.reset().clone()
        ms, dones, reloc = 0.0, 0, 0
        k = min(args.scripted, n) if task == 2 else 0
        prev_blk = env.get_state()["aux"][:3].clone() if task == 2 else None
        for t in range(args.steps + 10):
            a = torch.rand((n, 6), device=dev, generator=g) * 2 - 1
            if k:  # servo the first k envs' end effector onto their block
                q = obs[:k, :6].T.contiguous()
                z = torch.zeros_like(q)
                base = env.forward_dynamics(q, z, q)[3][:3]
                J = torch.zeros((3, 6, k), device=dev)
                for j in range(6):
                    qp = q.clone(); qp[j] += 1e-3
                    J[:, j] = (env.forward_dynamics(qp, z, qp)[3][:3] - base) / 1e-3
                err = env.get_state()["block"][:3, :k] - base
                a[:k] = torch.clamp(torch.einsum("cjk,ck->kj", J, err) * 400.0, -1, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = env.step(a)
            e1.record()
            obs = r.obs.clone()
            if t >= 10:
                e1.synchronize()
                ms += e0.elapsed_time(e1)
                dones += int((r.terminated | r.truncated).sum())
                if task == 2:
                    blk = env.get_state()["aux"][:3]
                    reloc += int(((blk - prev_blk).abs().sum(0) > 0).sum())
                    prev_blk = blk.clone()
            elif task == 2:
                prev_blk = env.get_state()["aux"][:3].clone()
        sec = ms * 1e-3
        out[cfg] = {"envs": n, "steps": args.steps, "ms_per_step": ms / args.steps, "env_steps_per_s": n * args.steps / sec,
                    "episode_resets_per_s": dones / sec, "resets": dones}
        if task == 2:
            out[cfg].update({"scripted_envs": k, "relocations": reloc, "relocations_per_s": reloc / sec})
        env.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
