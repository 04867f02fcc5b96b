#!/usr/bin/env python
"""Times the fused PPO kernels (include/so100_ppo.h) at BASELINE config 5's geometry: 65 536 envs x 32 steps, 8 minibatches.

    python tools/bench_ppo.py [--od 8] [--reps 20]

Algorithmic FLOPs per sample: forward 2*(od*64 + 64*64 + 64*nout) per tower, backward 2x that for the two 64-wide
layers' weight and input gradients (no input gradient for layer 1) -> see flops_per_sample().
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops_per_sample(od, train=True):
    f = 0
    for nout in (6, 1):
        fwd = 2 * (od * 64 + 64 * 64 + 64 * nout)
        bwd = 2 * (od * 64) + 2 * 2 * (64 * 64) + 2 * 2 * (64 * nout)  # dW1; dW2 + dH1; dW3 + dH2
        f += fwd + (bwd if train else 0)
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--od", type=int, default=8)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    from so100_mujoco_rl_b200 import _native
    from so100_mujoco_rl_b200.ppo import MlpPolicy, pack_params
    L, check = _native.lib(), _native.check
    od, S = args.od, args.envs * args.steps
    mb = S // args.minibatches
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    P = pack_params(MlpPolicy(od, 6)).to(dev)
    obs, act = torch.randn(S, od, device=dev), torch.randn(S, 6, device=dev)
    logp, adv, ret = torch.randn(S, device=dev) * 0.1 - 5.5, torch.randn(S, device=dev), torch.randn(S, device=dev)
    perm = torch.randperm(S, device=dev)
    ws = torch.zeros(int(L.so100_ppo_workspace_floats(od)), device=dev)
    grad, loss = torch.zeros(P.numel(), device=dev), torch.zeros(3, device=dev)
    m, v, step = torch.zeros_like(P), torch.zeros_like(P), torch.zeros(1, device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / reps

    k = [0]

    def grad_step():
        idx = perm[(k[0] % args.minibatches) * mb:(k[0] % args.minibatches + 1) * mb]
        k[0] += 1
        check(L.so100_ppo_grad(od, P.data_ptr(), obs.data_ptr(), act.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), idx.data_ptr(),
                               mb, 0.2, 0.5, 0.0, 1, ws.data_ptr(), grad.data_ptr(), loss.data_ptr(), st))

    def adam_step():
        check(L.so100_ppo_adam(P.numel(), P.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(), 1.0, 0.5, 3e-4, 0.9, 0.999, 1e-5, st))

    n = args.envs
    a_raw, a_clip, lp, val = torch.zeros(n, 6, device=dev), torch.zeros(n, 6, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    oc = torch.zeros(n, od, device=dev)

    def act_step():
        check(L.so100_ppo_act(od, P.data_ptr(), obs.data_ptr(), n, 0, 0, k[0], 0, a_raw.data_ptr(), a_clip.data_ptr(), lp.data_ptr(), val.data_ptr(), oc.data_ptr(), st))

    g_ms, a_ms, act_ms = timed(grad_step, args.reps), timed(adam_step, args.reps), timed(act_step, args.reps)
    out = {"obs_dim": od, "minibatch": mb, "grad_ms": g_ms, "adam_ms": a_ms, "act_ms": act_ms,
           "grad_tflops": mb * flops_per_sample(od) / (g_ms * 1e-3) / 1e12, "act_tflops": n * flops_per_sample(od, False) / (act_ms * 1e-3) / 1e12,
           "update_samples_per_s": mb / ((g_ms + a_ms) * 1e-3), "flops_per_sample_train": flops_per_sample(od)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
