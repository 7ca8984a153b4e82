#!/usr/bin/env python3
"""Small reproductions for the long-segment kernel: every space-free run of more than 256 bytes of a fuzz case,
encoded alone (and with a little context), compared with the oracle."""
import os, random, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import textgen
import wordpiece_b200
from _oracle import Oracle

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1140
rng = random.Random(seed)
n = int(10 ** rng.uniform(1.0, 6.4))
kw = dict(invalid_rate=rng.choice([0.0, 0.0, 0.0005, 0.01, 0.2]), long_run_rate=rng.choice([0.0, 0.0, 0.02, 0.1]),
          long_tokens=rng.choice([0, 0, 5, 40]))
text, vocab = textgen.case(seed, n, **kw)
if rng.random() < 0.15:
    vocab = [t for t in vocab if t != b"[UNK]"]
o = Oracle(vocab)
v = wordpiece_b200.Vocab(vocab, device=0)
words = [(m.start(), m.end()) for m in re.finditer(rb"[^ \n\t\r]{257,}", text)]
print(f"seed {seed}: {len(words)} long runs", flush=True)
bad = 0
for a, b in words[:400]:
    for pre, post in ((0, 0), (64, 64), (3000, 3000)):
        lo, hi = max(0, a - pre), min(len(text), b + post)
        while lo > 0 and text[lo - 1:lo] not in (b" ", b"\n"):
            lo -= 1
        while hi < len(text) and text[hi:hi + 1] not in (b" ", b"\n"):
            hi += 1
        t = text[lo:hi]
        exp = o.encode(t)
        got = v.encode(t)
        if not np.array_equal(exp, got):
            bad += 1
            if bad <= 6:
                m = min(len(exp), len(got))
                d = np.nonzero(exp[:m] != got[:m])[0]
                k = int(d[0]) if d.size else m
                print(f"MISMATCH run [{a},{b}) len {b - a} ctx {pre}: sizes {len(exp)} vs {len(got)} first at {k}: "
                      f"{exp[max(0,k-2):k+5].tolist()} vs {got[max(0,k-2):k+5].tolist()}", flush=True)
                print("   run head", text[a:a + 40], flush=True)
print("bad", bad)
