"""Shared parity cases.

``REFERENCE_GOLDEN`` restates, as data, the 28 known-answer checks of the
reference's own tests (tests/tests.cpp:137-217); ``UNK`` is -1 because none of
those vocabularies has an ``[UNK]`` line (tests/tests.cpp:17).
``REFERENCE_DIFFERENTIAL`` are its two hand-written linear==fast checks.
``QUIRKS`` are the behaviours of SURVEY.md Appendix A.3 (pinned by running the
compiled reference; expected ids recorded in tests/golden/quirks.json).
``fuzz_case`` draws the hostile tiny cases of Appendix A.5.
"""
from __future__ import annotations

import random

UNK = -1

REFERENCE_GOLDEN = [
    # testSimple  tests.cpp:139-161
    ("abcdef", ["bcde", "ac", "def", "bc", "bcdef", "a"], [UNK]),
    ("abcdef", ["bcde", "ac", "def", "bc", "##bcdef", "a"], [5, 4]),
    ("   aaaa  ", ["aa", "##aa"], [0, 1]),
    ("   aaaa  ", ["aa"], [UNK]),
    ("aaaa", ["aaaa"], [0]),
    ("aaaa", ["##aaaa"], [UNK]),
    ("aaaa", ["aaaa", "##aaaa", "##aaa", "##aa", "##a"], [0]),
    ("aaaa", ["##aaa", "aaaa", "##aa", "##a"], [1]),
    ("aaaa", ["aaa", "##aa", "##a", "##aaa"], [0, 2]),
    ("aaaa", ["aa", "a", "##aa"], [0, 2]),
    ("aaaa", ["aa", "a", "##aaa"], [UNK]),
    ("aaaa", ["aa", "##a"], [0, 1, 1]),
    ("abcdef", ["##def", "abc"], [1, 0]),
    ("abcdef", ["##bcde", "##ac", "##def", "##bc", "##bcdef", "a", "##a"], [5, 4]),
    ("abcdef", ["##bcdd", "##ac", "##def", "##bc", "##bcdff", "a"], [5, 3, 2]),
    ("djzhoyuhmcij", ["d", "##j", "##z", "##h", "##o", "##y", "##u", "##m", "##c", "##i", "##d"],
     [0, 1, 2, 3, 4, 5, 6, 3, 7, 8, 9, 1]),
    # testPunctuation  tests.cpp:165-167
    ("self-made", ["self", "made", "-", "##-", "##made"], [0, 2, 1]),
    ("self, made", ["self", "made", ",", "##,", "##made"], [0, 2, 1]),
    ("self  , made", ["self", "made", ",", "##,", "##made"], [0, 2, 1]),
    # testNonSplitted  tests.cpp:171-175
    ("abc", ["a", "abd"], [UNK]),
    ("abc a abc abd", ["a", "abd"], [UNK, 0, UNK, 1]),
    ("abcdef", ["bcde", "ac", "def", "bc", "bcdef", "##a", "##b", "##c", "##d"], [UNK]),
    # testMaxMatch  tests.cpp:180-184
    ("abcdef", ["a", "##bcdef", "ab", "##c", "##d", "##e", "##f"], [2, 3, 4, 5, 6]),
    ("abcdef abc abcd", ["abcd", "def", "abc"], [UNK, 2, 0]),
    # testUtf8  tests.cpp:209-216
    ("привет мир", ["привет", "мир"], [0, 1]),
    ("привет мир", ["при", "##вет", "мир"], [0, 1, 2]),
    ("токенизация это круто",
     ["ток", "крут", "это", "##за", "##ция", "ция"],
     [UNK, 2, UNK]),
    ("токенизация это круто",
     ["ток", "крут", "это", "##за", "##ени", "##о",
      "##ция", "ция"],
     [0, 4, 3, 6, 2, 1, 5]),
]

REFERENCE_DIFFERENTIAL = [
    # tests.cpp:138
    ("aaaa", ["aaaa", "aaa", "aa", "a"]),
    # tests.cpp:186-205 — note the missing comma in the source makes "##d##f" one token
    ("djzhoyuhmcijprfwrssuhvgzw",
     ["##c", "d", "##d##f", "##g", "##h", "##hv", "##i", "##j", "##m", "##o", "##p", "##r", "##s", "##u", "##uh",
      "##w", "##y", "##z"]),
]

ZH = "中"      # 中
WEN = "文"     # 文
ZI = "字"      # 字
HAN = "漢"     # 漢
KA = "か"      # か
NA = "な"      # な
A_HIRA = "あ"  # あ
MARU = "。"    # 。 (CJK full stop: NOT punctuation for the reference)
ASTRAL = "\U00020000"  # CJK ext B
COMPAT = "豈"  # U+8C48; U+F900 is its compatibility form
COMPAT_F900 = "豈"

# (name, text bytes, vocab) — SURVEY.md Appendix A.3; expected ids come from the compiled reference.
QUIRKS = [
    ("han_own_word", (ZH + "abc").encode(), [ZH, "abc", "[UNK]"]),
    ("han_oov_swallows_run", (ZH + "abc").encode(), ["abc", "[UNK]"]),
    ("han_fused_window", (ZH + "abc").encode(), [ZH, "abc", ZH + "abc", "[UNK]"]),
    ("han_fused_then_suffix", (ZH + "abcd").encode(), [ZH, "abc", ZH + "ab", "##cd", "##c", "[UNK]"]),
    ("han_fused_then_miss", (ZH + "abcz q").encode(), [ZH, "abc", ZH + "ab", "##c", "q", "[UNK]"]),
    ("kana_swallowed", (HAN + ZI + KA + NA).encode(), [HAN, KA + NA, "##" + KA + NA, "[UNK]"]),
    ("cjk_punct_not_punct", (ZH + MARU + WEN).encode(), [ZH, WEN, MARU, ZH + MARU, "[UNK]"]),
    ("punct_window_is_one", b"a-x", ["a", "-", "x", "-x", "[UNK]"]),
    ("token_spanning_punct_dead", b"self-made", ["self", "-", "made", "self-made", "[UNK]"]),
    ("token_with_space_dead", b"a b", ["a", "b", "a b", "[UNK]"]),
    ("duplicate_last_wins", b"aa", ["aa", "[UNK]", "aa"]),
    ("longer_than_max_len", b"abcdefgh xy", ["ab", "##cd", "xy", "[UNK]"]),
    ("unk_rolls_back", b"abcz abc", ["a", "##b", "##c", "[UNK]"]),
    ("specials_never_match", b"[CLS] a", ["[CLS]", "[", "]", "CLS", "a", "[UNK]"]),
    ("invalid_bytes_vanish", b"ab\xff\xfecd \xe2\x82 x", ["abcd", "x", "[UNK]"]),
    ("u2581_is_space", "a▁b".encode(), ["a", "b", "##b", "[UNK]"]),
    ("nbsp_is_ordinary", "a b".encode(), ["a", "b", "##b", "[UNK]"]),
    ("latin1_and_general_punct", "a·b—c".encode(), ["a", "b", "c", "·", "—", "[UNK]"]),
    ("no_unk_line", b"zz", ["a"]),
    ("last_unk_line_wins", b"z", ["[UNK]", "a", "[UNK]"]),
    ("triple_sharp_is_dead_suffix", b"a# #", ["a", "#", "###", "[UNK]"]),
    ("nul_is_ordinary", b"a\x00b", ["a", "b", "##b", "[UNK]"]),
    ("overlong_surrogate_range", b"a\xc0\x80b \xed\xa0\x80c \xf4\x90\x80\x80d \xc1\xbfe",
     ["ab", "c", "d", "e", "[UNK]"]),
    ("truncated_tail", b"ab \xe4\xb8", ["ab", "[UNK]"]),
    ("malformed_token_skipped", b"a -- b", ["a", "b", "--", "-", "[UNK]"]),
    ("astral_han", ("a" + ASTRAL + "b").encode(), ["a", "b", ASTRAL, "[UNK]"]),
    ("compat_han", (COMPAT_F900 + "x").encode(), [COMPAT_F900, "x", COMPAT_F900 + "x", "[UNK]"]),
    ("spaces_only_then_word", b" \t\n\r\x0b\x0c a", ["a", "[UNK]"]),
    ("crlf_tokens", b"a\r\nb", ["a", "b", "[UNK]"]),
    ("max_len_one_han", (ZH + "ab").encode(), ["a", "##b", "[UNK]"]),
    ("max_len_one_han_b", (ZH + "ab").encode(), ["a", "b", "##b", "z"]),
    ("suffix_special_lookalike", b"a[x]", ["a", "##[x]", "[", "]", "x", "[UNK]"]),
    ("sharp_word", b"## #a", ["#", "a", "##a", "[UNK]"]),
    ("zero_width_space_ordinary", "a​b c".encode(), ["a", "c", "a​b", "[UNK]"]),
    ("ideographic_space_ordinary", "a　b".encode(), ["a", "b", "##b", "[UNK]"]),
    ("general_punct_range_edges", "a‐b›c※d‏e".encode(),
     ["a", "b", "c", "d", "e", "‐", "›", "※", "##※d", "##‏e", "[UNK]"]),
    ("han_range_edges", "㏿a 㐀a 䶿a ䷀a 鿿a ꀀa".encode(),
     ["a", "㏿", "㐀", "䶿", "䷀", "鿿", "ꀀ", "㏿a", "䷀a", "ꀀa", "[UNK]"]),
    ("long_word_pieces", b"abcdefghijklmnopqrstuvwxyz0123456789abcdefghijklmnopqrstuvwxyz",
     ["abcdefghijklmnopqrstuvwxyz0123456789", "##abcdefghijklm", "##nopqrstuvwxyz", "[UNK]"]),
]

_ALPHABET = ["a", "b", "c", "A", "я", "ж", A_HIRA, KA, ZH, WEN, ZI, MARU, "-", ",", "#", "[", "]",
             "·"]
_TEXT_EXTRA = ["—", " ", " ", " ", "\n", "▁", " ", "", ASTRAL, "é"]
_INVALID = [b"\xff", b"\xc0", b"\x80", b"\xe2\x82", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xc1\xbf", b"\xf8",
            b"\xe4\xb8", b"\xf0\x9f"]


def fuzz_case(rng: random.Random, max_syms: int = 24):
    """One hostile tiny (text_bytes, vocab_list_of_bytes) case."""
    max_tok = rng.choice([1, 1, 2, 3, 4, 6])
    n_tok = rng.randint(1, 14)
    vocab = []
    for _ in range(n_tok):
        w = "".join(rng.choice(_ALPHABET) for _ in range(rng.randint(1, max_tok)))
        if rng.random() < 0.4:
            w = "##" + w
        if rng.random() < 0.05:
            w = "[" + w + "]"
        b = w.encode()
        if rng.random() < 0.03:
            b += rng.choice(_INVALID)
        vocab.append(b)
    if rng.random() < 0.7:
        vocab.insert(rng.randint(0, len(vocab)), b"[UNK]")
    if rng.random() < 0.2:
        vocab.append(rng.choice(vocab))
    n_sym = rng.randint(0, max_syms)
    parts = []
    for _ in range(n_sym):
        r = rng.random()
        if r < 0.08:
            parts.append(rng.choice(_INVALID))
        elif r < 0.35:
            parts.append(rng.choice(_TEXT_EXTRA).encode())
        else:
            parts.append(rng.choice(_ALPHABET).encode())
    return b"".join(parts), vocab


def random_split_case(rng: random.Random, text_len: int, parts: int, positive: bool):
    """tests.cpp:99-135,219-246 — random [a-z] string, vocab = a random cut of it."""
    s = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(text_len))
    borders = {len(s)}
    while len(borders) < parts:
        borders.add(rng.randint(1, len(s) - 1))
    out = set()
    start = 0
    for b in sorted(borders):
        if start == 0:
            out.add(s[start:b])
        out.add("##" + s[start:b])
        start = b
    vocab = sorted(out)
    if not positive:
        vocab = vocab[1:]
    return s, vocab
