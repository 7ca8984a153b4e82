#!/usr/bin/env python3
"""Why staging copies use streaming stores: the H2D rate of a pinned buffer right after the CPU wrote it with
ordinary (cached) stores, against the same buffer once the lines have left the CPU caches.
    python tools/h2d_dirty_cache.py [--mib 8]"""
import argparse
import time

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=8)
a = ap.parse_args()
n = a.mib << 20
src = torch.from_numpy(np.random.default_rng(1).integers(0, 255, n, dtype=np.uint8))
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
evict_a = torch.empty(512 << 20, dtype=torch.uint8)
evict_b = torch.empty(512 << 20, dtype=torch.uint8)


def h2d():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.copy_(pin, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for threads in (1, 8):
    torch.set_num_threads(threads)
    h2d()
    dirty, clean = [], []
    for _ in range(7):
        pin.copy_(src)  # cached stores, `threads` threads
        dirty.append(h2d())
        pin.copy_(src)
        evict_b.copy_(evict_a)  # 1 GiB of other traffic: the buffer's lines leave the caches
        time.sleep(0.01)
        clean.append(h2d())
    d, c = sorted(dirty)[len(dirty) // 2], sorted(clean)[len(clean) // 2]
    print(f"{a.mib} MiB pinned -> device, written by {threads} thread(s): right after the write {d:.3f} ms = {n / d / 1e6:.1f} GB/s; "
          f"after the lines left the caches {c:.3f} ms = {n / c / 1e6:.1f} GB/s")
