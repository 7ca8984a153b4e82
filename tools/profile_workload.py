#!/usr/bin/env python
"""One workload, device-resident, a few encode calls: the thing to put under ncu.

    python tools/profile_workload.py <en|ru|ja|zh|adv|dirty|open> [--mib 128] [--reps 3]

Prints one JSON line (GB/s, per-kernel ms from the library's events).  Under
`ncu --metrics gpu__time_duration.sum` the launch list splits K2 from the long-segment kernel; under
`ncu --set full -k regex:wp_ -s <skip> -c <n>` it gives the per-line tables (tools/ncu_lines.py)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import wordpiece_b200  # noqa: E402
from wordpiece_b200 import synth  # noqa: E402

MIB = 1 << 20
SEEDS = {"en": 2, "ru": 31, "ja": 32, "zh": 33, "adv": 5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--mib", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    base = {"dirty": "en", "open": "en"}.get(a.workload, a.workload)
    g = synth.generator(base)
    text = g.generate(a.mib * MIB, seed=SEEDS.get(base, 21) if a.workload == base else 21)
    if a.workload == "dirty":
        text = synth.dirty_web(text, seed=22)
    elif a.workload == "open":
        text = synth.open_vocabulary(text, seed=23)
    v = wordpiece_b200.Vocab(g.spec.vocab, device=0)
    d_text = torch.from_numpy(text).cuda()
    d_ids = torch.empty(text.size, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    _, n_ids = v.encode_device(d_text, d_ids)
    st = v.stats()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        v.encode_device_async(d_text, d_ids, d_cnt)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    v.set_kernel_timing(True)
    v.encode_device_async(d_text, d_ids, d_cnt)
    k_ms, n_ranges = v.last_kernel_ms()
    v.set_kernel_timing(False)
    med = float(np.median(times))
    print(json.dumps({"workload": a.workload, "mib": a.mib, "ms": med, "gb_per_s": text.size / (med * 1e-3) / 1e9,
                      "ids": int(n_ids), "kernel_ms": dict(zip(["split", "match+long", "scatter"], [round(x, 4) for x in k_ms])),
                      "ranges": n_ranges, "dirty_tiles": st.dirty_tiles, "long_segments": st.long_segments,
                      "memo_hits": st.memo_hits}))
    v.close()


if __name__ == "__main__":
    main()
