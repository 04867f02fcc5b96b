#!/usr/bin/env python
"""Host link bandwidth as the e2e path sees it: pinned <-> device cudaMemcpyAsync of the per-step payload sizes."""
import json
import torch

dev = torch.device("cuda", 0)
out = {}
for name, nbytes in (("actions_1.5MiB", 65536 * 6 * 4), ("obs_3.75MiB", 65536 * 15 * 4), ("64MiB", 64 << 20)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for direction in ("h2d", "d2h"):
        best = 0.0
        for rep in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            b.record(); b.synchronize()
            if rep >= 2:
                best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        out[f"{direction}_{name}_GBps"] = round(best, 2)
print(json.dumps(out))
