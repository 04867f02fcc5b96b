#!/usr/bin/env python
"""Trains the GPU PPO on a batched so100 env and writes the learning curve (gpurun_out/ppo_<env>.json).

    python tools/train_ppo.py --env Env01 --num-envs 4096 --samples 30000000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        tools/train_ppo.py --env Env05 --num-envs 65536 --samples 2e9          # BASELINE config 5: env-sharded, data-parallel

Multi-GPU: one process per GPU, `--num-envs` envs on EACH rank (weak scaling), global env ids rank*num_envs...;
gradients are averaged with one flat NCCL all-reduce per minibatch; the env step path has no collective.
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
    ap.add_argument("--flags", type=int, default=0, help="SO100_FLAG_* bits for the env (16 = arm-floor contact)")
    ap.add_argument("--tag", default="", help="suffix of the output file name")
    ap.add_argument("--learner", default="fused", choices=["fused", "torch"], help="fused = include/so100_ppo.h kernels; torch = the PyTorch reference learner")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from so100_mujoco_rl_b200.batched_env import BatchedSo100Env
    from so100_mujoco_rl_b200.ppo import PPO, FusedPPO, PPOConfig
    from so100_mujoco_rl_b200.sharding import dist_env
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = BatchedSo100Env(args.env, args.num_envs, device=local_rank, seed=args.seed, env_offset=rank * args.num_envs,
                          max_episode_steps=args.max_episode_steps or None, flags=args.flags)
    cfg = PPOConfig(n_steps=args.n_steps, n_minibatches=args.minibatches, n_epochs=args.epochs, seed=args.seed)
    algo = FusedPPO(env, cfg, env_offset=rank * args.num_envs) if args.learner == "fused" else PPO(env, cfg)
    hist = []

    def cb(rec):
        hist.append(rec)
        if rank == 0 and (rec["iter"] % 20 == 0 or rec["iter"] == 1):
            print(json.dumps(rec), flush=True)
    t0 = time.time()
    st = algo.learn(int(args.samples), log_every=0, callback=cb)
    wall = time.time() - t0
    graphed = getattr(algo, "_graph", None) is not None
    if world > 1:
        # a captured CUDA graph holds NCCL kernels: drop it and drain the device before tearing the group down
        # (destroying the communicator under a live graph can hang at exit)
        algo._graph = None
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)
    out = {"env": args.env, "learner": args.learner, "num_envs": args.num_envs, "n_gpus": world, "cuda_graph_update": graphed, "n_steps": args.n_steps, "samples": st.samples, "wall_s": wall,
           "rollout_s": st.rollout_s, "update_s": st.update_s, "samples_per_s": st.samples / wall,
           "rollout_env_steps_per_s": st.samples / st.rollout_s, "history": hist[:: max(1, len(hist) // 200)] + hist[-1:],
           "kernel_variant": env.kernel_variant, "stats": env.stats(), "env_flags": args.flags}
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, f"ppo_{args.env}_{args.learner}_s{args.seed}" + (f"_{world}gpu" if world > 1 else "") + args.tag + ".json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: out[k] for k in ("n_gpus", "samples", "wall_s", "rollout_s", "update_s", "samples_per_s", "cuda_graph_update")})
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
