"""Byte-range sharding of a text across GPUs (BASELINE.json configs[3]).

The encode path shards with no exchange step: the reference's serial state is
reset at every ``is_space`` code point (fast.cpp:89-91, and its own chunking cuts
there, fast.cpp:113-115), so contiguous byte ranges whose cuts sit right after a
space byte sequence are encoded independently and the id arrays concatenated.
Global id offsets are an exclusive scan over the per-shard id counts.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_ASCII_SPACES = (0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20)


def _is_space_at(text: np.ndarray, i: int) -> int:
    """Length of the reference ``is_space`` byte sequence starting at i (0 if none): utf8.cpp:10-12."""
    b = int(text[i])
    if b in _ASCII_SPACES:
        return 1
    if b == 0xE2 and i + 2 < text.size and int(text[i + 1]) == 0x96 and int(text[i + 2]) == 0x81:
        return 3  # U+2581
    return 0


def plan_shards(text: np.ndarray, n_shards: int, search_limit: int = 1 << 26) -> List[Tuple[int, int]]:
    """Split ``text`` (uint8 array) into ``n_shards`` contiguous ranges of near-equal size.

    Each cut is moved forward to just after the next space byte sequence.  If no
    space occurs within ``search_limit`` bytes (space-free CJK corpora) the cut is
    moved to the start of the next ASCII punctuation byte instead, which is also a
    safe start.  Ranges may be empty when the text is tiny.
    """
    n = int(text.size)
    cuts = [0]
    for k in range(1, n_shards):
        pos = max(cuts[-1], (n * k) // n_shards)
        end = min(n, pos + search_limit)
        cut = None
        i = pos
        while i < end:
            ln = _is_space_at(text, i)
            if ln:
                cut = i + ln
                break
            i += 1
        if cut is None:
            i = pos
            while i < end:
                b = int(text[i])
                if b < 0x80 and (0x21 <= b <= 0x2F or 0x3A <= b <= 0x40 or 0x5B <= b <= 0x60 or 0x7B <= b <= 0x7E):
                    cut = i
                    break
                i += 1
        cuts.append(n if cut is None else cut)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


def global_offsets(counts: Sequence[int]) -> List[int]:
    """Exclusive scan of per-shard id counts -> offset of each shard's ids in the global array."""
    out, run = [], 0
    for c in counts:
        out.append(run)
        run += int(c)
    return out
