#!/usr/bin/env python
"""Host-link ceiling of the e2e path: how many env steps per second the link alone could carry.

One env step moves 24 B of actions host->device and (obs_dim * 4 + 6) B of obs / reward / flags device->host.  This
tool times, with nothing but copy engines (pinned host memory, no kernel), per rank:
  * each direction alone at the per-step payload sizes, and
  * BOTH directions at once, back to back without any dependency (two streams): the floor of a perfectly pipelined
    host path, `ceiling_env_steps_per_s` = envs / that time.
Run it alone (1 GPU) or under torchrun with N ranks at once: on a multi-GPU box the ranks share the host's memory and
root complexes, and bench.py's 8-GPU e2e number has to be read against the CONCURRENT ceiling.

    python tools/pcie_bw.py [--envs 65536] [--obs-dim 15]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw.py
"""
import argparse
import json
import os

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--obs-dim", type=int, default=15)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, od = args.envs, args.obs_dim
    in_b, out_b = n * 24, n * (od * 4 + 6)
    h_in, h_out = torch.empty(in_b, dtype=torch.uint8).pin_memory(), torch.empty(out_b, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(in_b, dtype=torch.uint8, device=dev), torch.empty(out_b, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        for _ in range(10):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            fn()
        s1.synchronize(); s2.synchronize()
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / args.reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
    if rank == 0:
        print(json.dumps({
            "ranks": world, "envs_per_rank": n, "h2d_bytes": in_b, "d2h_bytes": out_b,
            "h2d_alone_ms": t_in, "h2d_alone_GBps": in_b / t_in / 1e6, "d2h_alone_ms": t_out, "d2h_alone_GBps": out_b / t_out / 1e6,
            "both_directions_ms": t_both, "ceiling_env_steps_per_s_per_rank": n / (t_both * 1e-3),
            "ceiling_env_steps_per_s_all_ranks": n * world / (t_both * 1e-3),
            "note": "copy engines only, both directions streaming without dependencies; max over ranks when ranks > 1"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
