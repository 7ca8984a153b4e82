#!/usr/bin/env python3
"""PCIe ceiling for the host-buffer entry point: 1 GiB host->device and 1.16 GB device->host (the e2e
step of bench.py's English workload), alone and concurrently, pinned memory, one B200."""
import json
import time

import torch

n_in = 1 << 30
n_out = 1155161204
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


a, b, c = run(True, False), run(False, True), run(True, True)
print(json.dumps({"h2d_alone_ms": a * 1e3, "h2d_GBs": n_in / a / 1e9, "d2h_alone_ms": b * 1e3, "d2h_GBs": n_out / b / 1e9,
                  "both_ms": c * 1e3, "text_GBs_at_ceiling": n_in / c / 1e9}))
