"""TEST-ONLY byte-domain model of the kernel's decomposition.

Runs the same plan as wp_encode.cu — drop invalid bytes, find safe starts, match
each segment independently with longest-match queries against the TABLE IMAGE
(``wp_debug_longest_match``, the host mirror of the device probe) — in plain
Python, so that the vocabulary table and the segment logic can be checked against
the oracle without a GPU.  Never used by the product.
"""
from __future__ import annotations

from typing import List

from _oracle import Oracle

SPACE, PUNCT, HAN, OTHER = 1, 2, 3, 0


def _cls(cp: int) -> int:
    L = Oracle.lib()
    if L.wpo_is_space(cp):
        return SPACE
    if L.wpo_is_punct(cp):
        return PUNCT
    if L.wpo_is_han(cp):
        return HAN
    return OTHER


def model_encode(vocab_handle, text: bytes) -> List[int]:
    """vocab_handle: wordpiece_b200.Vocab (host-only is fine)."""
    cps, _ = Oracle.decode_utf8(text)
    chars = [chr(c).encode("utf-8") for c in cps]          # canonical bytes per char == clean text
    cls = [_cls(c) for c in cps]
    clean = b"".join(chars)
    off = [0]
    for ch in chars:
        off.append(off[-1] + len(ch))
    n = len(cps)
    unk = vocab_handle.unk_id
    swallow = vocab_handle.max_len >= 2
    out: List[int] = []

    def longest(i: int, j: int, kind: int):
        """longest token of `kind` that is a prefix of chars[i:j] -> (n_bytes, id)"""
        return vocab_handle.debug_longest_match(clean[off[i]:off[j]], kind)

    def byte_to_char(i: int, nbytes: int) -> int:
        target = off[i] + nbytes
        k = i
        while off[k] < target:
            k += 1
        assert off[k] == target, "match ended inside a character"
        return k

    i = 0
    while i < n:
        if cls[i] == SPACE:
            i += 1
            continue
        prev = cls[i - 1] if i > 0 else SPACE
        # safe start?  (inside a segment we never get here)
        assert cls[i] in (PUNCT, HAN) or prev in (SPACE, PUNCT), "model bug: not a safe start"
        if cls[i] == PUNCT:
            k, tid = longest(i, i + 1, 0)
            out.append(tid if k == len(chars[i]) else unk)
            i += 1
            continue
        e = i + 1
        while e < n and cls[e] == OTHER:
            e += 1
        seg: List[int] = []
        p = i
        kind = 0
        word_first = 0
        done = False
        if cls[i] == HAN:
            k, tid = longest(p, e, 0)
            if k == 0:
                seg.append(unk)
                if swallow:
                    done = True
                else:
                    word_first = 1
                    p += 1
            else:
                seg.append(tid)
                p = byte_to_char(p, k)
                if k == len(chars[i]):
                    word_first = 1
                else:
                    kind = 1
        while not done and p < e:
            k, tid = longest(p, e, kind)
            if k == 0:
                del seg[word_first:]
                seg.append(unk)
                break
            seg.append(tid)
            p = byte_to_char(p, k)
            kind = 1
        out.extend(seg)
        i = e
    return out


def word_hash_py(word: bytes, slots_log2: int) -> int:
    """wp_table.h word_hash: home slot of a word (<= 16 bytes) in a word table of 2**slots_log2 slots."""
    b = word + b"\0" * (16 - len(word))
    k = [int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(4)]
    m = 0xFFFFFFFF
    h = (k[0] * 0x9E3779B1 + k[1] * 0x85EBCA77 + k[2] * 0xC2B2AE3D + k[3] * 0x27D4EB2F + len(word) * 0x165667B1) & m
    h ^= h >> 15
    return ((h * 0x2C1B3C6D) & m) >> (32 - slots_log2)
