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
