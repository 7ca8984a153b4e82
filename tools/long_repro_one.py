import os, random, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import textgen
import wordpiece_b200
from _oracle import Oracle
seed, a, b, ctx = 1140, 61333, 61931, 3000
rng = random.Random(seed)
n = int(10 ** rng.uniform(1.0, 6.4))
kw = dict(invalid_rate=rng.choice([0.0, 0.0, 0.0005, 0.01, 0.2]), long_run_rate=rng.choice([0.0, 0.0, 0.02, 0.1]), long_tokens=rng.choice([0, 0, 5, 40]))
text, vocab = textgen.case(seed, n, **kw)
o = Oracle(vocab)
v = wordpiece_b200.Vocab(vocab, device=0)
lo, hi = max(0, a - ctx), min(len(text), b + ctx)
while lo > 0 and text[lo - 1:lo] not in (b" ", b"\n"): lo -= 1
while hi < len(text) and text[hi:hi + 1] not in (b" ", b"\n"): hi += 1
t = text[lo:hi]
print("text len", len(t), "run at", a - lo, "..", b - lo, flush=True)
print("long runs in t:", [(m.start(), m.end()) for m in re.finditer(rb"[^ \n\t\r]{257,}", t)], flush=True)
for rep in range(3):
    exp = o.encode(t); got = v.encode(t)
    print("rep", rep, "equal", np.array_equal(exp, got), len(exp), len(got), flush=True)
