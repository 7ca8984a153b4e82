"""Regenerates the golden fixtures from the UNMODIFIED reference compiled into
oracle/_ref/ (oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

quirks.json   expected ids of tests/cases.py:QUIRKS   
mixed.json    FNV-1a-64 + length + head of the reference's ids for the seeded
              texts of tests/textgen.py (multi-tile, invalid bytes, long runs)
fuzz.json     2 000 hostile tiny cases (text/vocab hex + reference ids)
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import textgen  # noqa: E402
from _oracle import Ref, fnv1a64, in_reference_domain  # noqa: E402

assert Ref.available(), "build oracle/_ref first: make -C oracle ref"

quirks = {}
for name, text, vocab in cases.QUIRKS:
    assert in_reference_domain(text, vocab), name
    quirks[name] = {"text_hex": text.hex(), "vocab": vocab, "ids": Ref.encode(text, vocab).tolist()}
json.dump(quirks, open(os.path.join(HERE, "quirks.json"), "w"), indent=1, ensure_ascii=True)

mixed = {}
for seed, n, kw in [(11, 5000, {}), (15, 40000, {}), (16, 300000, {}), (22, 30000, dict(invalid_rate=0.05)),
                    (24, 60000, dict(invalid_rate=0.5)), (31, 50000, dict(long_run_rate=0.08, long_tokens=40))]:
    text, vocab = textgen.case(seed, n, **kw)
    ids = Ref.encode(text, vocab)
    mixed[f"seed{seed}"] = {"seed": seed, "n_bytes": n, "kw": kw, "n_ids": int(ids.size), "fnv1a64": f"{fnv1a64(ids):016x}",
                            "head": ids[:32].tolist()}
json.dump(mixed, open(os.path.join(HERE, "mixed.json"), "w"), indent=1)

rng = random.Random(777)
fuzz = []
while len(fuzz) < 2000:
    text, vocab = cases.fuzz_case(rng)
    if not in_reference_domain(text, vocab):
        continue
    fuzz.append({"t": text.hex(), "v": [t.hex() for t in vocab], "ids": Ref.encode(text, vocab).tolist()})
json.dump(fuzz, open(os.path.join(HERE, "fuzz.json"), "w"), separators=(",", ":"))
print("golden written:", len(quirks), "quirks,", len(mixed), "mixed,", len(fuzz), "fuzz")
