#!/usr/bin/env python3
"""Encoder::encodeBatch (C++) into a fresh std::vector<int> with and without the huge-page advice on its storage
(WORDPIECE_B200_HUGEPAGES=0 / unset), interleaved, through lib/dropin_bench.   python tools/batch_hugepage_ab.py"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wordpiece_b200 import synth  # noqa: E402

g = synth.generator("en")
text = g.generate(48 << 20, seed=2)
exe = os.path.join(ROOT, "wordpiece_b200", "lib", "dropin_bench")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    tf, vf = os.path.join(d, "text.txt"), os.path.join(d, "vocab.txt")
    text.tofile(tf)
    with open(vf, "wb") as f:
        f.write(b"\n".join(t if isinstance(t, bytes) else t.encode() for t in g.spec.vocab) + b"\n")
    for env in [dict(), dict(WORDPIECE_B200_HUGEPAGES="0")] * 2:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([exe, tf, vf, "1", "10000", "4096"], capture_output=True, text=True, env=e, timeout=120)
        try:
            b = json.loads(r.stdout.strip().splitlines()[-1])["batch"]
            print(json.dumps({"advice": "off" if env else "on (default)", "fresh_vector_ms": round(1e3 * b["fresh_vector_best_seconds"], 2),
                              "reused_vector_ms": round(1e3 * b["reused_vector_best_seconds"], 2), "texts": b["texts"], "bytes": b["bytes"]}), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"env": env, "error": repr(ex), "stderr": r.stderr[-300:]}), flush=True)
