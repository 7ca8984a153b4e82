#!/usr/bin/env python3
"""Where does the time of a 4 KiB call go?  Host enqueue time, device time between events, full synchronous call.
    python tools/small_call.py [--bytes 4096] [--calls 2000]        (under ncu: --calls 20 for the launch list)"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import wordpiece_b200  # noqa: E402
from wordpiece_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=4096)
ap.add_argument("--calls", type=int, default=2000)
a = ap.parse_args()
g = synth.generator("en")
text = g.generate(8 << 20, seed=2)
v = wordpiece_b200.Vocab(g.spec.vocab, device=0)
rng = np.random.default_rng(3)
starts = rng.integers(0, text.size - 2 * a.bytes, size=64)
d_texts = []
for s in starts:
    s = int(s)
    while text[s - 1] != 0x20:
        s += 1
    e = s + a.bytes
    while text[e - 1] != 0x20:
        e -= 1
    d_texts.append(torch.from_numpy(text[s:e].copy()).cuda())
d_ids = torch.empty(a.bytes, dtype=torch.int32, device="cuda")
d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for t in d_texts[:20]:
    v.encode_device(t, d_ids)
sync, enq, dev = [], [], []
for i in range(a.calls):
    t = d_texts[i % len(d_texts)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v.encode_device(t, d_ids)
    sync.append(time.perf_counter() - t0)
for i in range(a.calls):
    t = d_texts[i % len(d_texts)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    t0 = time.perf_counter()
    v.encode_device_async(t, d_ids, d_cnt)
    enq.append(time.perf_counter() - t0)
    e1.record()
    torch.cuda.synchronize()
    dev.append(e0.elapsed_time(e1) * 1e-3)


def p(x, q):
    return sorted(x)[int(q * (len(x) - 1))] * 1e6


print(f"{a.bytes}-byte texts, {a.calls} calls: synchronous encode_device p50 {p(sync, .5):.1f} us p99 {p(sync, .99):.1f} us; "
      f"encode_device_async host enqueue p50 {p(enq, .5):.1f} us; device time between events p50 {p(dev, .5):.1f} us")
