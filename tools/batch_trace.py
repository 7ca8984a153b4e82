#!/usr/bin/env python3
"""Milestones of one wp_encode_batch call (10 000 x 4 KiB) on stderr: WORDPIECE_B200_TRACE=1 python tools/batch_trace.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import wordpiece_b200  # noqa: E402
from wordpiece_b200 import synth  # noqa: E402

g = synth.generator("en")
text = g.generate(64 << 20, seed=2)
rng = np.random.default_rng(1)
slices = []
for start in rng.integers(0, text.size - 8192, size=10000):
    s = int(start)
    while text[s - 1] != 0x20:
        s += 1
    e = s + 4096
    while text[e - 1] != 0x20:
        e -= 1
    slices.append(text[s:e].tobytes())
total = sum(len(b) for b in slices)
v = wordpiece_b200.Vocab(g.spec.vocab, device=0)
prepared = v.batch_pointers(slices)
out = torch.empty(total, dtype=torch.int32, pin_memory=True).numpy()
offs = np.zeros(len(slices) + 1, np.uint64)
for rep in range(4):
    t0 = time.perf_counter()
    v.encode_batch(slices, out=out, offsets=offs, prepared=prepared)
    print(f"rep {rep}: {(time.perf_counter() - t0) * 1e3:.3f} ms for {total} bytes", file=sys.stderr)

# where does the device time of a batch go?  one part (6 MiB of the same texts), per-kernel events, next to a plain
# device-resident encode of the same packed bytes
small = slices[:1500]
v.set_kernel_timing(True)
ids, offs2 = v.encode_batch(small)
k_ms, n_ranges = v.last_kernel_ms()
print(f"kernel ms of a one-part batch ({sum(len(b) for b in small)} bytes, {n_ranges} ranges): split {k_ms[0]:.3f} match {k_ms[1]:.3f} scatter {k_ms[2]:.3f}", file=sys.stderr)
packed = b"".join(b + b" " for b in small)
d_text = torch.frombuffer(bytearray(packed), dtype=torch.uint8).cuda()
d_ids = torch.empty(len(packed), dtype=torch.int32, device="cuda")
d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
v.encode_device_async(d_text, d_ids, d_cnt)
k_ms, n_ranges = v.last_kernel_ms()
print(f"kernel ms of a plain encode of the same bytes ({n_ranges} ranges): split {k_ms[0]:.3f} match {k_ms[1]:.3f} scatter {k_ms[2]:.3f}", file=sys.stderr)
v.set_kernel_timing(False)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v.encode_device(d_text, d_ids)
    print(f"plain device-resident encode of the packed bytes: {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr)
