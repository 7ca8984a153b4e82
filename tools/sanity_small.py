#!/usr/bin/env python
"""Small exact-size workload for compute-sanitizer (memcheck / racecheck): every lane of the encode path —
clean tiles, tiles with invalid UTF-8, segments that leave their window, several ranges, misaligned and
exact-size device buffers, the host-buffer pipeline — checked against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import textgen
import wordpiece_b200
from _oracle import Oracle

tile = wordpiece_b200.tile_bytes()
for seed, n, kw in [(1, 3 * tile + 17, {}), (2, 2 * tile + 5, dict(invalid_rate=0.05)),
                    (3, 5 * tile - 3, dict(long_run_rate=0.1, long_tokens=10)), (4, 100, {}), (5, tile, {})]:
    text, vocab = textgen.case(seed, n, **kw)
    exp = Oracle(vocab).encode(text)
    v = wordpiece_b200.Vocab(vocab, device=0)
    assert np.array_equal(v.encode(text), exp), seed
    d = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda().clone()          # exact-size allocation
    ids, cnt = v.encode_device(d, torch.empty(len(exp), dtype=torch.int32, device="cuda"))  # exact capacity
    assert cnt == len(exp) and np.array_equal(ids[:cnt].cpu().numpy(), exp), seed
    sub = d[3:]
    e2 = Oracle(vocab).encode(text[3:])
    ids2, c2 = v.encode_device(sub)
    assert c2 == len(e2) and np.array_equal(ids2[:c2].cpu().numpy(), e2), seed
    os.environ["WORDPIECE_B200_RANGE_BYTES"] = str(2 * tile)
    assert np.array_equal(v.encode(text), exp), seed
    del os.environ["WORDPIECE_B200_RANGE_BYTES"]
    os.environ["WORDPIECE_B200_PIPE_CHUNK"] = "4096"
    out = np.zeros(len(exp) + 1, np.int32)
    assert v.encode_into(text, out) == len(exp) and np.array_equal(out[:-1], exp), seed
    del os.environ["WORDPIECE_B200_PIPE_CHUNK"]
    v.close()
print("sanity ok")
