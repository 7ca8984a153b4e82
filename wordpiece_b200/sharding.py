"""Byte-range sharding of a text across GPUs (BASELINE.json configs[3]) — thin ctypes layer.

The logic lives in the C ABI (``wp_plan_shards``, ``wp_encode_sharded[_gather]``,
csrc/wp_capi.cu): the encode path shards with no exchange step — the reference's
serial state is reset at every ``is_space`` code point (fast.cpp:89-91, and its own
chunking cuts there, fast.cpp:113-115) — so contiguous byte ranges cut at safe
starts are encoded independently and the id arrays concatenated.  Global id offsets
are an exclusive scan over the per-shard id counts.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._capi import Vocab, _buffer_address, _check, load_library


class _ShardStruct(C.Structure):
    _fields_ = [("begin", C.c_size_t), ("end", C.c_size_t), ("n_ids", C.c_uint64), ("id_offset", C.c_uint64),
                ("device", C.c_int), ("encode_ms", C.c_float)]


@dataclass
class Shard:
    begin: int
    end: int
    n_ids: int
    id_offset: int
    device: int
    encode_ms: float


def plan_shards(text, n_shards: int) -> List[int]:
    """``wp_plan_shards``: the ``n_shards + 1`` cut offsets (first 0, last len) of near-equal shards, each cut
    moved forward to the next safe cut (after a space; for space-free CJK text at punctuation / Han chars)."""
    L = load_library()
    addr, n, keep = _buffer_address(text)
    cuts = (C.c_size_t * (n_shards + 1))()
    got = int(L.wp_plan_shards(addr, n, n_shards, cuts))
    del keep
    if got != n_shards + 1:
        raise ValueError("wp_plan_shards: bad arguments")
    return [int(c) for c in cuts]


def next_safe_cut(text, pos: int) -> int:
    """``wp_next_safe_cut``: the first safe cut at or after ``pos`` (len(text) if none)."""
    L = load_library()
    addr, n, keep = _buffer_address(text)
    cut = int(L.wp_next_safe_cut(addr, n, pos))
    del keep
    return cut


def shard_ranges(text, n_shards: int) -> List[Tuple[int, int]]:
    cuts = plan_shards(text, n_shards)
    return list(zip(cuts[:-1], cuts[1:]))


def global_offsets(counts: Sequence[int]) -> List[int]:
    """Exclusive scan of per-shard id counts -> offset of each shard's ids in the global array."""
    out, run = [], 0
    for c in counts:
        out.append(run)
        run += int(c)
    return out


def _handles(vocabs: Sequence[Vocab]):
    return (C.c_void_p * len(vocabs))(*[v._h for v in vocabs])


def encode_sharded(vocabs: Sequence[Vocab], text, out: np.ndarray) -> Tuple[int, List[Shard]]:
    """``wp_encode_sharded``: host text -> host ids over ``len(vocabs)`` GPUs (one handle per device)."""
    assert out.dtype == np.int32 and out.flags.c_contiguous
    L = load_library()
    addr, n, keep = _buffer_address(text)
    cnt = C.c_size_t()
    sh = (_ShardStruct * len(vocabs))()
    _check(L.wp_encode_sharded(_handles(vocabs), len(vocabs), addr, n, out.ctypes.data, out.size, C.byref(cnt), sh))
    del keep
    return int(cnt.value), [Shard(s.begin, s.end, s.n_ids, s.id_offset, s.device, s.encode_ms) for s in sh]


def encode_sharded_gather(vocabs: Sequence[Vocab], text, d_ids, gather_index: int = 0):
    """``wp_encode_sharded_gather``: ids gathered peer-to-peer into the int32 CUDA tensor ``d_ids`` on the device
    of ``vocabs[gather_index]``.  Returns (count, shards, gather_ms)."""
    L = load_library()
    addr, n, keep = _buffer_address(text)
    cnt = C.c_size_t()
    ms = C.c_float()
    sh = (_ShardStruct * len(vocabs))()
    _check(L.wp_encode_sharded_gather(_handles(vocabs), len(vocabs), addr, n, gather_index, d_ids.data_ptr(),
                                      d_ids.numel(), C.byref(cnt), sh, C.byref(ms)))
    del keep
    return (int(cnt.value), [Shard(s.begin, s.end, s.n_ids, s.id_offset, s.device, s.encode_ms) for s in sh],
            float(ms.value))
