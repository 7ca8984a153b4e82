"""Synthetic vocabularies and corpora for the BASELINE.json configurations.

No network in the build or GPU containers, so the workloads are synthetic but
shaped like the reference's benchmark inputs (bert-base-cased vocabulary +
Wikipedia text, reference README.md:47, tests/speed_test.py:126-151):

=========  ===============================================================
config     vocabulary / text
=========  ===============================================================
``en``     "bert-cased-29k" (28 996 lines) / Zipf English-like ASCII text
``ru``     "mbert-120k" (119 547 lines) / Cyrillic words, ASCII separators
``ja``     "mbert-120k" / Han + kana runs without spaces, some OOV Han
``zh``     "mbert-120k" / Han runs, one token per char, some OOV Han
``adv``    "long-m100" (en vocab + 2 000 tokens of 60-100 chars) / high-UNK
=========  ===============================================================

Vocabularies and lexicons are built in numpy/Python (seeded, < 10 s); the text
itself is produced by ``lib/libwp_synth.so`` (csrc/wp_synth.c) in independent
1 MiB blocks, so a block-aligned shard of a 10 GB corpus can be generated on its
own rank and equals that range of the whole corpus.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB = os.path.join(_HERE, "lib", "libwp_synth.so")
BLOCK = 1 << 20

_ONSETS = ["", "b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w", "z", "br", "ch",
           "cl", "cr", "dr", "fl", "fr", "gr", "pl", "pr", "qu", "sc", "sh", "sl", "sp", "st", "str", "th", "tr", "wh"]
_NUCLEI = ["a", "e", "i", "o", "u", "a", "e", "i", "o", "ai", "ea", "ee", "ie", "io", "oo", "ou", "y"]
_CODAS = ["", "", "", "n", "r", "s", "t", "l", "d", "m", "ng", "nt", "st", "ck", "ll", "ss", "rd", "nd", "ct", "x"]
_CYR_CONS = list("бвгджзклмнпрстфхцчшщ")
_CYR_VOW = list("аеиоуыэюя")
_HIRA = [chr(c) for c in range(0x3041, 0x3097)]
_KATA = [chr(c) for c in range(0x30A1, 0x30FB)]


@dataclass
class Spec:
    """A corpus: weighted items + weighted separators (+ capitalisation probability)."""
    name: str
    vocab: List[bytes]
    items: List[bytes]
    item_weights: np.ndarray
    seps: List[bytes]
    sep_weights: np.ndarray
    cap_prob: float = 0.0


def _alias_tables(weights: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Vose's alias method; thresholds scaled to 2^32."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    p = w / w.sum() * n
    prob = np.zeros(n, np.float64)
    alias = np.arange(n, dtype=np.uint32)
    small = [i for i in range(n) if p[i] < 1.0]
    large = [i for i in range(n) if p[i] >= 1.0]
    p = p.copy()
    while small and large:
        s = small.pop()
        l = large.pop()
        prob[s] = p[s]
        alias[s] = l
        p[l] = p[l] + p[s] - 1.0
        (small if p[l] < 1.0 else large).append(l)
    for i in large + small:
        prob[i] = 1.0
    thr = np.minimum(prob * 4294967296.0, 4294967295.0).astype(np.uint32)
    return thr, alias


def _pack(strings: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(strings) + 1, np.uint32)
    off[1:] = np.cumsum([len(s) for s in strings], dtype=np.uint64).astype(np.uint32)
    data = np.frombuffer(b"".join(strings) or b"\0", dtype=np.uint8).copy()
    return data, off


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB):
            raise RuntimeError(f"{SYNTH_LIB} is missing: run `make -C wordpiece_b200/csrc`")
        L = C.CDLL(SYNTH_LIB)
        vp = C.c_void_p
        L.wp_synth_fill.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_uint64, vp, vp, vp, vp, C.c_uint32, vp, vp, vp,
                                    vp, C.c_uint32, C.c_uint32, C.c_int]
        L.wp_synth_fill.restype = C.c_int
        _lib = L
    return _lib


class Generator:
    """Binds a Spec to the C generator."""

    def __init__(self, spec: Spec):
        self.spec = spec
        self._ib, self._io = _pack(spec.items)
        self._ip, self._ia = _alias_tables(spec.item_weights)
        self._sb, self._so = _pack(spec.seps)
        self._sp, self._sa = _alias_tables(spec.sep_weights)
        self._cap = int(min(max(spec.cap_prob, 0.0), 1.0) * 4294967295.0)

    def fill(self, out: np.ndarray, seed: int, first_block: int = 0, n_threads: int = 0) -> np.ndarray:
        """Fill the uint8 array ``out`` with corpus bytes [first_block MiB, +len(out))."""
        assert out.dtype == np.uint8 and out.flags.c_contiguous
        if n_threads <= 0:
            n_threads = min(32, os.cpu_count() or 1)
        rc = _load().wp_synth_fill(out.ctypes.data, out.size, first_block, seed & 0xFFFFFFFFFFFFFFFF,
                                   self._ib.ctypes.data, self._io.ctypes.data, self._ip.ctypes.data,
                                   self._ia.ctypes.data, len(self.spec.items), self._sb.ctypes.data,
                                   self._so.ctypes.data, self._sp.ctypes.data, self._sa.ctypes.data,
                                   len(self.spec.seps), self._cap, n_threads)
        if rc != 0:
            raise RuntimeError(f"wp_synth_fill failed ({rc})")
        return out

    def generate(self, n_bytes: int, seed: int, first_block: int = 0, n_threads: int = 0) -> np.ndarray:
        return self.fill(np.empty(n_bytes, np.uint8), seed, first_block, n_threads)


# ------------------------------------------------------------------ lexicons

def _latin_word(rng: np.random.Generator) -> str:
    n_syl = int(min(5, max(1, round(rng.lognormal(0.55, 0.45)))))
    w = ""
    for _ in range(n_syl):
        w += _ONSETS[rng.integers(len(_ONSETS))] + _NUCLEI[rng.integers(len(_NUCLEI))] + _CODAS[rng.integers(len(_CODAS))]
    return w[:20]


def _cyr_word(rng: np.random.Generator) -> str:
    n_syl = int(min(6, max(1, round(rng.lognormal(0.95, 0.4)))))
    w = ""
    for _ in range(n_syl):
        w += _CYR_CONS[rng.integers(len(_CYR_CONS))] + _CYR_VOW[rng.integers(len(_CYR_VOW))]
        if rng.random() < 0.35:
            w += _CYR_CONS[rng.integers(len(_CYR_CONS))]
    return w[:18]


def _unique_words(make, rng, n: int) -> List[str]:
    seen, out = set(), []
    guard = 0
    while len(out) < n and guard < 60 * n:
        guard += 1
        w = make(rng)
        if w and w not in seen:
            seen.add(w)
            out.append(w)
    # rank: shorter words tend to be more frequent
    keys = np.array([len(w) for w in out], dtype=np.float64) + rng.normal(0, 1.6, len(out))
    order = np.argsort(keys, kind="stable")
    return [out[i] for i in order]


def _zipf(n: int, s: float = 1.0) -> np.ndarray:
    return 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)


def _specials(n_unused_head: int = 99, filler_to: int = 1000) -> List[str]:
    v = ["[PAD]"] + [f"[unused{i}]" for i in range(1, n_unused_head + 1)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    i = n_unused_head + 1
    while len(v) < filler_to:
        v.append(f"[unused{i}]")
        i += 1
    return v


_ASCII_PUNCT = [chr(c) for c in range(0x21, 0x7F) if not chr(c).isalnum()]
_ASCII_ALNUM = [chr(c) for c in range(0x21, 0x7F) if chr(c).isalnum()]
_EXTRA_SINGLES = ([chr(c) for c in range(0xC0, 0x100) if c not in (0xD7, 0xF7)] + [chr(c) for c in range(0x391, 0x3CA) if c != 0x3A2]
                  + [chr(c) for c in range(0x410, 0x450)])
_ACCENTED = list("éèüöäñçåøß")


@functools.lru_cache(maxsize=None)
def _english_parts(vocab_size: int = 28996, n_types: int = 200_000):
    rng = np.random.default_rng(20240517)
    lex = _unique_words(_latin_word, rng, n_types)
    vocab: List[str] = _specials()
    vocab += _ASCII_PUNCT + _ASCII_ALNUM + ["##" + c for c in _ASCII_ALNUM]
    han = [chr(c) for c in rng.choice(np.arange(0x4E00, 0x9FFF), size=500, replace=False)]
    vocab += _EXTRA_SINGLES + han + _HIRA[:60] + _KATA[:60]          # word-initial only: "##é" etc. are absent
    n_suffix = 6000
    n_prefix = vocab_size - len(vocab) - n_suffix
    # word-initial tokens: the most frequent words, a capitalised form for the top of the list
    n_cap = 2500
    prefix = lex[: n_prefix - n_cap] + [w.capitalize() for w in lex[:n_cap]]
    # continuation tokens: frequent word endings (suffixes of lexicon words, weighted by rank)
    counts = {}
    for r, w in enumerate(lex[:60000]):
        wt = 1.0 / (r + 1)
        for k in (1, 2, 3, 4, 5, 6):
            if len(w) > k:
                tail = w[-k:]
                counts[tail] = counts.get(tail, 0.0) + wt
                mid = w[1:1 + k]
                counts[mid] = counts.get(mid, 0.0) + 0.3 * wt
    singles = set(_ASCII_ALNUM)
    tails = [t for t, _ in sorted(counts.items(), key=lambda kv: -kv[1]) if t not in singles][:n_suffix]
    vocab += prefix + ["##" + t for t in tails]
    seen, uniq = set(), []
    for t in vocab:
        if t not in seen:
            seen.add(t)
            uniq.append(t)
    i = 10_000
    while len(uniq) < vocab_size:  # top up after de-duplication
        t = f"[unused{i}]"
        i += 1
        if t not in seen:
            seen.add(t)
            uniq.append(t)
    return uniq[:vocab_size], lex


def english_vocab() -> List[bytes]:
    """"bert-cased-29k": 28 996 unique lines, [UNK] at index 100, max token length ~20."""
    return [t.encode("utf-8") for t in _english_parts()[0]]


def english_spec() -> Spec:
    vocab, lex = _english_parts()
    rng = np.random.default_rng(1)
    items = list(lex)
    weights = _zipf(len(lex), 1.0)
    total = weights.sum()
    # ~2.5 % of running words carry a letter that has no "##" form -> whole-word UNK
    acc = []
    for w in lex[200:6200:3]:
        pos = 1 + int(rng.integers(max(1, len(w) - 1)))
        acc.append(w[:pos] + _ACCENTED[int(rng.integers(len(_ACCENTED)))] + w[pos:])
    acc_w = np.full(len(acc), 0.025 * total / len(acc))
    # ~2 % digit groups
    nums = [str(int(x)) for x in rng.integers(0, 100000, size=3000)] + [str(y) for y in range(1900, 2030)]
    num_w = np.full(len(nums), 0.02 * total / len(nums))
    items = items + acc + nums
    weights = np.concatenate([weights, acc_w, num_w])
    seps = [" ", ", ", ". ", "; ", ": ", "! ", "? ", " (", ") ", " - ", "\n", ".\n", "  ", "'s ", "\" "]
    sep_w = np.array([83, 5, 3.5, 0.6, 0.6, 0.3, 0.4, 0.5, 0.5, 0.4, 2.5, 1.5, 1.0, 0.5, 0.2])
    return Spec("en", [t.encode() for t in vocab], [w.encode() for w in items], weights, [s.encode() for s in seps],
                sep_w, cap_prob=0.10)


@functools.lru_cache(maxsize=None)
def _mbert_parts(vocab_size: int = 119_547):
    rng = np.random.default_rng(3)
    en_vocab, en_lex = _english_parts()
    cyr_lex = _unique_words(_cyr_word, rng, 150_000)
    han_all = np.arange(0x4E00, 0x9FFF)
    han_in = [chr(c) for c in rng.choice(han_all, size=6000, replace=False)]
    han_set = set(han_in)
    han_oov = [chr(c) for c in han_all if chr(c) not in han_set][:3000]
    kana = _HIRA + _KATA
    vocab: List[str] = _specials()
    vocab += _ASCII_PUNCT + _ASCII_ALNUM + ["##" + c for c in _ASCII_ALNUM]
    cyr_letters = [chr(c) for c in range(0x410, 0x450)] + ["ё", "Ё"]
    vocab += _EXTRA_SINGLES + cyr_letters + ["##" + c for c in cyr_letters]
    vocab += han_in + ["##" + c for c in han_in[:1500]] + kana + ["##" + k for k in kana] + ["。", "、", "「", "」"]
    # kana n-grams (2-4), word-initial and continuation
    grams = set()
    while len(grams) < 9000:
        n = int(rng.integers(2, 5))
        src = _HIRA if rng.random() < 0.7 else _KATA
        grams.add("".join(src[int(rng.integers(len(src)))] for _ in range(n)))
    grams = sorted(grams)
    vocab += grams[:6000] + ["##" + g for g in grams[3000:9000]]
    # Han compounds fused with kana (exercise the CJK-start window)
    fused = set()
    while len(fused) < 2000:
        fused.add(han_in[int(rng.integers(len(han_in)))] + grams[int(rng.integers(len(grams)))][:2])
    vocab += sorted(fused)
    # Cyrillic words and endings
    counts = {}
    for r, w in enumerate(cyr_lex[:50000]):
        wt = 1.0 / (r + 1)
        for k in (2, 3, 4, 5):
            if len(w) > k:
                counts[w[-k:]] = counts.get(w[-k:], 0.0) + wt
                counts[w[2:2 + k]] = counts.get(w[2:2 + k], 0.0) + 0.3 * wt
    cyr_tails = [t for t, _ in sorted(counts.items(), key=lambda kv: -kv[1])][:12000]
    vocab += cyr_lex[:30000] + [w.capitalize() for w in cyr_lex[:3000]] + ["##" + t for t in cyr_tails]
    # Latin as in the English vocabulary
    vocab += [t for t in en_vocab if not t.startswith("[")]
    seen, uniq = set(), []
    for t in vocab:
        if t not in seen:
            seen.add(t)
            uniq.append(t)
    k = 0
    while len(uniq) < vocab_size:
        w = en_lex[40000 + k]
        k += 1
        if w not in seen:
            seen.add(w)
            uniq.append(w)
    return uniq[:vocab_size], cyr_lex, han_in, han_oov, grams, sorted(fused)


def mbert_vocab() -> List[bytes]:
    """"mbert-120k": 119 547 unique lines over Latin, Cyrillic, kana and ~6 000 Han chars."""
    return [t.encode("utf-8") for t in _mbert_parts()[0]]


def russian_spec() -> Spec:
    vocab, cyr_lex, *_ = _mbert_parts()
    weights = _zipf(len(cyr_lex), 1.0)
    seps = [" ", ", ", ". ", "; ", ": ", " - ", "\n", ".\n", " («", "») ", "  "]
    sep_w = np.array([82, 6, 4, 0.5, 0.5, 1.0, 2.5, 1.5, 0.5, 0.5, 1.0])
    return Spec("ru", [t.encode() for t in vocab], [w.encode() for w in cyr_lex], weights, [s.encode() for s in seps],
                sep_w, cap_prob=0.0)


def japanese_spec() -> Spec:
    vocab, _, han_in, han_oov, grams, fused = _mbert_parts()
    items = han_in + han_oov[:1500] + grams + fused
    w = np.concatenate([
        _zipf(len(han_in), 0.9) * 1.0,
        np.full(1500, 0.05 * _zipf(len(han_in), 0.9).sum() / 0.95 / 1500 * 0.45),  # ~5 % of Han chars are OOV
        _zipf(len(grams), 0.8) * 1.2,
        _zipf(len(fused), 0.8) * 0.15,
    ])
    seps = ["", "。", "、", " ", ",", "\n", "「", "」"]
    sep_w = np.array([88, 3.5, 4.5, 1.0, 0.5, 1.5, 0.5, 0.5])
    return Spec("ja", [t.encode() for t in vocab], [s.encode() for s in items], w, [s.encode() for s in seps], sep_w)


def chinese_spec() -> Spec:
    vocab, _, han_in, han_oov, _, _ = _mbert_parts()
    items = han_in + han_oov[:2000]
    base = _zipf(len(han_in), 0.95)
    w = np.concatenate([base, np.full(2000, 0.03 * base.sum() / 0.97 / 2000)])  # ~3 % OOV Han
    seps = ["", ",", ".", " ", "\n", "!", "?"]
    sep_w = np.array([93, 3.0, 1.5, 1.0, 1.0, 0.25, 0.25])
    return Spec("zh", [t.encode() for t in vocab], [s.encode() for s in items], w, [s.encode() for s in seps], sep_w)


def adversarial_vocab() -> List[bytes]:
    """"long-m100": the English vocabulary + 2 000 tokens of 60-100 chars (max_len = 100)."""
    vocab, lex = _english_parts()
    rng = np.random.default_rng(5)
    longs, seen = [], set(vocab)
    while len(longs) < 2000:
        n = int(rng.integers(60, 101))
        w = ""
        while len(w) < n:
            w += lex[int(rng.integers(50000))]
        w = w[:n]
        t = w if rng.random() < 0.5 else "##" + w
        if t not in seen:
            seen.add(t)
            longs.append(t)
    return [t.encode() for t in vocab + longs]


def adversarial_spec() -> Spec:
    """High-UNK text: half the words are random 20-120 char strings whose alphabet
    includes letters without a "##" form, half are ordinary English words."""
    en = english_spec()
    rng = np.random.default_rng(51)
    alphabet = list("abcdefghijklmnopqrstuvwxyz") * 3 + _ACCENTED
    pool = []
    for _ in range(40000):
        n = int(rng.integers(20, 121))
        pool.append("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=n)))
    total = en.item_weights.sum()
    items = en.items + [p.encode() for p in pool]
    weights = np.concatenate([en.item_weights, np.full(len(pool), total / len(pool))])
    return Spec("adv", adversarial_vocab(), items, weights, en.seps, en.sep_weights, cap_prob=0.05)


SPECS = {
    "en": english_spec,
    "ru": russian_spec,
    "ja": japanese_spec,
    "zh": chinese_spec,
    "adv": adversarial_spec,
}


@functools.lru_cache(maxsize=None)
def generator(name: str) -> Generator:
    return Generator(SPECS[name]())


def corpus(name: str, n_bytes: int, seed: int, first_block: int = 0, n_threads: int = 0) -> Tuple[np.ndarray, List[bytes]]:
    """(text uint8 array, vocab) of configuration ``name``."""
    g = generator(name)
    return g.generate(n_bytes, seed, first_block, n_threads), g.spec.vocab


def write_vocab_file(path: str, vocab: Sequence[bytes]) -> None:
    with open(path, "wb") as f:
        for t in vocab:
            f.write(t + b"\n")


# --------------------------------------------------------------- derived shapes

_B64 = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
_INVALID_BYTES = np.array([0xFF, 0xC0, 0x80, 0xFE, 0xF8, 0xBF, 0xC1, 0xE2], dtype=np.uint8)


def dirty_web(clean: np.ndarray, seed: int, invalid_rate: float = 0.01, long_token_rate: float = 0.005,
              long_min: int = 300, long_max: int = 4000) -> np.ndarray:
    """"Ordinary dirty web text" of the same size as ``clean`` (an ASCII corpus): ``long_token_rate`` of the
    tokens are space-free strings of ``long_min``..``long_max`` bytes (base64 blobs; a third of them URL-like,
    with a '/' every 20-60 bytes), and ``invalid_rate`` of the bytes are overwritten with bytes that are
    invalid UTF-8 where they stand (at 1 % every 4 KiB tile holds some)."""
    rng = np.random.default_rng(seed)
    n = clean.size
    mean_tokens = 1.0 / max(long_token_rate, 1e-9)
    parts, size, pos = [], 0, 0
    while size < n:
        take = int(rng.exponential(mean_tokens * 5.7)) + 1           # ~5.7 bytes per running token
        end = min(n, pos + take)
        while end < n and clean[end - 1] != 0x20:
            end += 1
        if end <= pos:
            pos = 0
            continue
        parts.append(clean[pos:end])
        size += end - pos
        pos = end if end < n else 0
        ln = int(rng.integers(long_min, long_max + 1))
        blob = _B64[rng.integers(0, _B64.size, ln)]
        if rng.random() < 0.33:
            k = 0
            while True:
                k += int(rng.integers(20, 61))
                if k >= ln:
                    break
                blob[k] = 0x2F
        parts.append(blob)
        parts.append(np.array([0x20], np.uint8))
        size += ln + 1
    out = np.concatenate(parts)[:n].copy()
    k = int(n * invalid_rate)
    if k:
        at = rng.integers(0, n, k)
        out[at] = _INVALID_BYTES[rng.integers(0, _INVALID_BYTES.size, k)]
    return out


def open_vocabulary(clean: np.ndarray, seed: int, hapax_rate: float = 0.03) -> np.ndarray:
    """An ASCII corpus with an OPEN vocabulary: ``hapax_rate`` of the running words are replaced by strings
    that occur only once (random alphanumerics of 3-14 chars: numbers, names, ids, typos), so that a cache
    of matched words cannot have seen them.  Same size as ``clean``."""
    rng = np.random.default_rng(seed)
    n = clean.size
    spaces = np.flatnonzero(clean == 0x20)
    pick = spaces[rng.random(spaces.size) < hapax_rate]
    out = clean.copy()
    lens = rng.integers(3, 15, pick.size)
    alnum = _B64
    for p, ln in zip(pick.tolist(), lens.tolist()):
        a = p + 1
        b = min(n, a + ln)
        if b < n:
            out[a:b] = alnum[rng.integers(0, alnum.size, b - a)]
            out[b] = 0x20
    return out
