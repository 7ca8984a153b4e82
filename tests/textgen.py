"""Seeded hostile text/vocab generators for the parity tests (small, pure Python)."""
from __future__ import annotations

import random
from typing import List, Tuple

import cases

_WORD_CHARS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJ0123456789"
_CYR = "абвгдежзиклмнопрстуфхцчшщыэюя"
_KANA = "あいうえおかきくけこさしすせそたちつてとなにぬねの"
_HAN = "中文字漢語日本人大小山川田"
_PUNCT = list(".,;:!?-()[]\"'#") + ["·", "—", "«", "»", "‐"]
_SPACES = [" ", " ", " ", " ", "\n", "\t", "  ", "▁", "\r\n"]
_ORDINARY_ODD = [" ", "​", "　", "é", "\x00", "\x7f", "。", "、"]


def mixed_vocab(rng: random.Random, n_words: int = 400, with_unk: bool = True, long_tokens: int = 0) -> List[bytes]:
    """A vocabulary over several scripts with prefix and ## tokens, dead tokens, duplicates."""
    vocab: List[str] = []
    if with_unk:
        vocab += ["[PAD]", "[UNK]", "[CLS]", "[SEP]"]
    singles = list(_WORD_CHARS) + list(_CYR) + list(_KANA) + list(_HAN[:8]) + _PUNCT[:12]
    for ch in singles:
        if rng.random() < 0.9:
            vocab.append(ch)
        if ch not in _PUNCT and rng.random() < 0.8:
            vocab.append("##" + ch)
    for _ in range(n_words):
        script = rng.choice([_WORD_CHARS[:26]] * 4 + [_CYR] * 2 + [_KANA])
        w = "".join(rng.choice(script) for _ in range(rng.randint(2, 9)))
        vocab.append(w if rng.random() < 0.6 else "##" + w)
    # fused Han tokens, dead tokens, specials-in-text, duplicates
    vocab += ["中文", "中abc", "字かな", "a b", "self-made", "-x", "...", "##...", "[x]", "日本"]
    for _ in range(long_tokens):
        w = "".join(rng.choice(_WORD_CHARS[:26]) for _ in range(rng.randint(23, 90)))
        vocab.append(w if rng.random() < 0.5 else "##" + w)
    rng.shuffle(vocab)
    if len(vocab) > 10:
        vocab.append(vocab[7])  # duplicate: last index wins
    seen, out = set(), []
    for t in vocab:
        out.append(t.encode("utf-8"))
        seen.add(t)
    return out


def mixed_text(rng: random.Random, n_bytes: int, vocab: List[bytes], invalid_rate: float = 0.0,
               long_run_rate: float = 0.0, max_run: int = 600) -> bytes:
    """Text drawn from vocab pieces, random words, CJK runs, punctuation, odd spaces, invalid bytes."""
    words = [t.decode("utf-8", "ignore").lstrip("#") for t in vocab if not t.startswith(b"[")]
    words = [w for w in words if w]
    parts: List[bytes] = []
    size = 0
    while size < n_bytes:
        r = rng.random()
        if r < 0.45:
            w = rng.choice(words)
            if rng.random() < 0.4:
                w += rng.choice(words)
            b = w.encode()
        elif r < 0.60:
            script = rng.choice([_WORD_CHARS, _CYR, _KANA])
            b = "".join(rng.choice(script) for _ in range(rng.randint(1, 12))).encode()
        elif r < 0.72:
            b = "".join(rng.choice(_HAN + _KANA) for _ in range(rng.randint(1, 10))).encode()
        elif r < 0.80:
            b = rng.choice(_PUNCT).encode() * rng.randint(1, 3)
        elif r < 0.84:
            b = rng.choice(_ORDINARY_ODD).encode()
        elif r < 0.84 + long_run_rate:
            ln = rng.randint(64, max_run)
            b = "".join(rng.choice(_WORD_CHARS[:26] + _CYR[:6]) for _ in range(ln)).encode()
        else:
            b = b""
        if invalid_rate and rng.random() < invalid_rate:
            junk = rng.choice(cases._INVALID)
            pos = rng.randint(0, len(b))
            b = b[:pos] + junk + b[pos:]
        sep = rng.choice(_SPACES).encode() if rng.random() < 0.8 else b""
        parts.append(b + sep)
        size += len(b) + len(sep)
    return b"".join(parts)[:n_bytes]


def case(seed: int, n_bytes: int, **kw) -> Tuple[bytes, List[bytes]]:
    rng = random.Random(seed)
    vocab = mixed_vocab(rng, long_tokens=kw.pop("long_tokens", 0))
    return mixed_text(rng, n_bytes, vocab, **kw), vocab
