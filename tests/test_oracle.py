"""CPU tests that PIN the oracle (oracle/wp_oracle.c) to the reference:
the reference's own golden vectors, fixtures recorded from the compiled
reference, and — when oracle/_ref is present — live differential fuzzing."""
from __future__ import annotations

import json
import os
import random

import numpy as np
import pytest

import cases
import textgen
from _oracle import EmptyVocabWord, Oracle, Ref, fnv1a64, in_reference_domain

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_golden_vectors():
    """tests/tests.cpp:137-217 (28 known-answer checks)."""
    assert len(cases.REFERENCE_GOLDEN) == 28
    for text, vocab, expected in cases.REFERENCE_GOLDEN:
        assert Oracle(vocab).encode(text).tolist() == expected, (text, vocab)


def test_recorded_quirks():
    golden = json.load(open(os.path.join(HERE, "golden", "quirks.json")))
    for name, text, vocab in cases.QUIRKS:
        assert text.hex() == golden[name]["text_hex"], f"{name}: fixture out of date"
        assert Oracle(vocab).encode(text).tolist() == golden[name]["ids"], name


def test_recorded_fuzz():
    fuzz = json.load(open(os.path.join(HERE, "golden", "fuzz.json")))
    assert len(fuzz) == 2000
    for c in fuzz:
        text = bytes.fromhex(c["t"])
        vocab = [bytes.fromhex(t) for t in c["v"]]
        assert Oracle(vocab).encode(text).tolist() == c["ids"]


def test_recorded_mixed_texts():
    mixed = json.load(open(os.path.join(HERE, "golden", "mixed.json")))
    for key, rec in mixed.items():
        text, vocab = textgen.case(rec["seed"], rec["n_bytes"], **rec["kw"])
        ids = Oracle(vocab).encode(text)
        assert ids.size == rec["n_ids"], key
        assert ids[:32].tolist() == rec["head"], key
        assert f"{fnv1a64(ids):016x}" == rec["fnv1a64"], key


def test_character_classes_appendix_c():
    L = Oracle.lib()
    spaces = {0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x2581}
    punct = set(range(0x21, 0x30)) | set(range(0x3A, 0x41)) | set(range(0x5B, 0x61)) | set(range(0x7B, 0x7F))
    punct |= {0xAB, 0xB7, 0xBB} | set(range(0x2010, 0x203B))
    for cp in list(range(0, 0x3100)) + [0x3400, 0x4DBF, 0x4DC0, 0x4DFF, 0x4E00, 0x9FFF, 0xA000, 0xF8FF, 0xF900, 0xFAFF,
                                        0xFB00, 0x1FFFF, 0x20000, 0x2A6DF, 0x2A6E0, 0x2A700, 0x2B73F, 0x2B740, 0x2B81F,
                                        0x2B820, 0x2CEAF, 0x2CEB0, 0x2F7FF, 0x2F800, 0x2FA1F, 0x2FA20, 0x10FFFF]:
        assert bool(L.wpo_is_space(cp)) == (cp in spaces), hex(cp)
        assert bool(L.wpo_is_punct(cp)) == (cp in punct), hex(cp)
    han_ranges = [(0x3400, 0x4DBF), (0x4E00, 0x9FFF), (0xF900, 0xFAFF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F),
                  (0x2B740, 0x2B81F), (0x2B820, 0x2CEAF), (0x2F800, 0x2FA1F)]
    for lo, hi in han_ranges:
        for cp in (lo - 1, lo, hi, hi + 1):
            inside = any(a <= cp <= b for a, b in han_ranges)
            assert bool(L.wpo_is_han(cp)) == inside, hex(cp)


def test_strict_utf8_decoder():
    dec = Oracle.decode_utf8
    assert dec(b"a\xc3\xa9\xe4\xb8\xad\xf0\xa0\x80\x80") == ([0x61, 0xE9, 0x4E2D, 0x20000], False)
    for bad in (b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xed\xa0\x80", b"\xf0\x80\x80\x80", b"\xf4\x90\x80\x80",
                b"\xf8\x88\x80\x80\x80", b"\x80", b"\xbf", b"\xe4\xb8", b"\xf0\x9f\x98", b"\xc3"):
        cps, invalid = dec(b"x" + bad + b"y")
        assert invalid and cps[0] == 0x78 and cps[-1] == 0x79, bad
        assert all(c < 0x80 for c in cps), bad  # nothing but the ASCII survives
    assert dec(b"\xef\xbf\xbf\xf4\x8f\xbf\xbf") == ([0xFFFF, 0x10FFFF], False)


def test_vocab_classification():
    o = Oracle(["a", "##b", "[CLS]", "--", "##..", "[", "[]", "[UNK]", "x[UNK]", "##[y]", "-", "##"[:2] + "#"])
    flags = [o.token_flags(i) for i in range(12)]
    assert flags == [1, 0, 3, 5, 4, 1, 5, 3, 1, 0, 1, 0]
    assert o.unk_id == 7 and o.max_len == 6
    with pytest.raises(EmptyVocabWord):
        Oracle(["a", "##"])
    with pytest.raises(EmptyVocabWord):
        Oracle([b"\xff\xfe"])


def test_outside_reference_domain_is_defined():
    # the reference divides by zero here (fast.cpp:45); the oracle continues with the natural reading
    assert Oracle(["a"]).encode(b"\xff\xfe").size == 0
    assert Oracle(["[UNK]", "[CLS]"]).encode(b"ab cd").tolist() == [0, 0]
    assert not in_reference_domain(b"\xff", ["a"]) and not in_reference_domain(b"a", ["[UNK]"])


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs the reference sources)")
def test_live_against_compiled_reference():
    for text, vocab, expected in cases.REFERENCE_GOLDEN:
        assert Ref.encode(text, vocab).tolist() == expected
    rng = random.Random(4242)
    n = 0
    for _ in range(4000):
        text, vocab = cases.fuzz_case(rng, max_syms=40)
        if not in_reference_domain(text, vocab):
            continue
        n += 1
        assert np.array_equal(Ref.encode(text, vocab), Oracle(vocab).encode(text)), (text, vocab)
    assert n > 3500
    for seed, nb, kw in [(101, 60000, {}), (102, 80000, dict(invalid_rate=0.03)),
                         (103, 50000, dict(long_run_rate=0.05, long_tokens=30))]:
        text, vocab = textgen.case(seed, nb, **kw)
        assert np.array_equal(Ref.encode(text, vocab), Oracle(vocab).encode(text)), seed


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs the reference sources)")
def test_reference_random_split_stress_small():
    """tests.cpp:219-246 shape; linear == fast == oracle inside the agreement domain."""
    rng = random.Random(17)
    for text_len in (10, 45, 120, 300):
        for parts in (2, 9, 40, 100):
            if parts > text_len:
                continue
            for positive in (True, False):
                s, vocab = cases.random_split_case(rng, text_len, parts, positive)
                f = Ref.encode(s, vocab)
                assert np.array_equal(f, Ref.encode(s, vocab, "linear"))
                assert np.array_equal(f, Oracle(vocab).encode(s))
