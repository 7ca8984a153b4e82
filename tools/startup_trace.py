#!/usr/bin/env python3
"""Where does a short process spend its time?  Writes BASELINE configs[0] (10 MB English, 29k vocabulary) to
/dev/shm and runs `runner fast <text> <vocab> 8` with WORDPIECE_B200_TRACE=1 (milestones on stderr), three times,
next to the reference's runner.  usage: python tools/startup_trace.py"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from wordpiece_b200 import synth  # noqa: E402

g = synth.generator("en")
text = g.generate(10_000_000, seed=2)
cut = text.size
while text[cut - 1] != 0x20:
    cut -= 1
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    tf, vf = os.path.join(d, "t.txt"), os.path.join(d, "v.txt")
    text[:cut].tofile(tf)
    with open(vf, "wb") as f:
        f.write(b"\n".join(t if isinstance(t, bytes) else t.encode() for t in g.spec.vocab) + b"\n")
    for tag, exe in (("ours", os.path.join(ROOT, "wordpiece_b200", "lib", "runner")),
                     ("reference", os.path.join(ROOT, "oracle", "_ref", "runner"))):
        if not os.path.exists(exe):
            continue
        for rep in range(3):
            env = dict(os.environ, WORDPIECE_B200_TRACE="1")
            t0 = time.perf_counter()
            r = subprocess.run([exe, "fast", tf, vf, "8"], capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            print(f"== {tag} run {rep}: {dt * 1e3:.1f} ms wall, rc {r.returncode}, {r.stdout.strip()}")
            if rep == 2:
                print(r.stderr.strip())
