"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  — oracle/_build/libwp_oracle.so, our plain-C restatement
  (oracle/wp_oracle.c) of reference src/fast.cpp:19-150.
* ``Ref``     — oracle/_ref/libwpref.so, the UNMODIFIED reference compiled by
  oracle/Makefile (absent if the recipe was never run; tests that need it skip).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libwp_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libwpref.so")
REF_RUNNER = os.path.join(ORACLE_DIR, "_ref", "runner")
REF_TESTS = os.path.join(ORACLE_DIR, "_ref", "tests")


def _as_bytes(x) -> bytes:
    return x.encode("utf-8") if isinstance(x, str) else bytes(x)


def _vocab_arrays(vocab: Sequence):
    toks = [_as_bytes(t) for t in vocab]
    n = len(toks)
    arr = (C.c_char_p * max(n, 1))(*toks) if n else (C.c_char_p * 1)()
    lens = (C.c_size_t * max(n, 1))(*[len(t) for t in toks]) if n else (C.c_size_t * 1)()
    return toks, arr, lens, n


def build_oracle() -> str:
    """Compile the C restatement if needed (gcc, < 1 s)."""
    src = os.path.join(ORACLE_DIR, "wp_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


class EmptyVocabWord(RuntimeError):
    pass


class Oracle:
    """The C restatement.  ``Oracle(vocab).encode(text) -> np.int32 array``."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(build_oracle())
            L.wpo_vocab_create.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t,
                                           C.POINTER(C.c_void_p)]
            L.wpo_vocab_create.restype = C.c_int
            L.wpo_vocab_free.argtypes = [C.c_void_p]
            L.wpo_vocab_unk_id.argtypes = [C.c_void_p]
            L.wpo_vocab_unk_id.restype = C.c_int32
            L.wpo_vocab_max_len.argtypes = [C.c_void_p]
            L.wpo_vocab_max_len.restype = C.c_size_t
            L.wpo_vocab_token_flags.argtypes = [C.c_void_p, C.c_size_t]
            L.wpo_vocab_token_flags.restype = C.c_int
            L.wpo_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_int32)),
                                     C.POINTER(C.c_size_t)]
            L.wpo_encode.restype = C.c_int
            L.wpo_free.argtypes = [C.c_void_p]
            L.wpo_decode_utf8.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
            L.wpo_decode_utf8.restype = C.c_size_t
            for f in ("wpo_is_space", "wpo_is_punct", "wpo_is_han", "wpo_is_spacing"):
                getattr(L, f).argtypes = [C.c_uint32]
                getattr(L, f).restype = C.c_int
            cls._lib = L
        return cls._lib

    def __init__(self, vocab: Sequence):
        L = self.lib()
        self._toks, arr, lens, n = _vocab_arrays(vocab)
        h = C.c_void_p()
        rc = L.wpo_vocab_create(arr, lens, n, C.byref(h))
        if rc == 1:
            raise EmptyVocabWord("Vocab word is empty")
        if rc != 0:
            raise MemoryError("wpo_vocab_create")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib().wpo_vocab_free(h)
            self._h = None

    @property
    def unk_id(self) -> int:
        return int(self.lib().wpo_vocab_unk_id(self._h))

    @property
    def max_len(self) -> int:
        return int(self.lib().wpo_vocab_max_len(self._h))

    def token_flags(self, i: int) -> int:
        return int(self.lib().wpo_vocab_token_flags(self._h, i))

    def encode(self, text) -> np.ndarray:
        L = self.lib()
        b = _as_bytes(text) if not isinstance(text, np.ndarray) else text.tobytes()
        ids = C.POINTER(C.c_int32)()
        n = C.c_size_t()
        rc = L.wpo_encode(self._h, b, len(b), C.byref(ids), C.byref(n))
        if rc != 0:
            raise MemoryError("wpo_encode")
        out = np.ctypeslib.as_array(ids, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int32)
        if n.value:
            L.wpo_free(ids)
        return out.astype(np.int32, copy=False)

    @classmethod
    def decode_utf8(cls, b: bytes):
        L = cls.lib()
        out = (C.c_uint32 * max(len(b), 1))()
        bad = C.c_int()
        n = L.wpo_decode_utf8(b, len(b), out, C.byref(bad))
        return list(out[:n]), bool(bad.value)


def oracle_encode(text, vocab) -> np.ndarray:
    return Oracle(vocab).encode(text)


def in_reference_domain(text, vocab) -> bool:
    """False where the unmodified reference crashes (SIGFPE at fast.cpp:45: the
    text decodes to zero code points, or no token is usable) or throws."""
    b = _as_bytes(text)
    try:
        o = Oracle(vocab)
    except EmptyVocabWord:
        return False
    if len(b) == 0:
        return True  # fast.cpp:145 returns {} before the division
    cps, _ = Oracle.decode_utf8(b)
    return len(cps) > 0 and o.max_len > 0


class Ref:
    """The unmodified reference through oracle/ref_capi.cpp."""

    _lib = None

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls, n_threads: int = 0):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.wpref_init_threads.argtypes = [C.c_size_t]
            L.wpref_init_threads.restype = C.c_size_t
            L.wpref_last_error.restype = C.c_char_p
            L.wpref_free.argtypes = [C.c_void_p]
            L.wpref_encode.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p),
                                       C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.POINTER(C.c_int)),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
            L.wpref_encode.restype = C.c_int
            L.wpref_encode_files.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(C.c_int)),
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
            L.wpref_encode_files.restype = C.c_int
            L.wpref_encode_external.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t,
                                                C.POINTER(C.c_double)]
            L.wpref_encode_external.restype = C.c_int
            L.wpref_decode.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t)]
            L.wpref_decode.restype = C.c_int
            cls.pool_threads = int(L.wpref_init_threads(n_threads))
            cls._lib = L
        return cls._lib

    @classmethod
    def encode(cls, text, vocab, algo: str = "fast", return_seconds: bool = False):
        L = cls.lib()
        b = _as_bytes(text) if not isinstance(text, np.ndarray) else text.tobytes()
        _, arr, lens, n = _vocab_arrays(vocab)
        ids = C.POINTER(C.c_int)()
        cnt = C.c_size_t()
        sec = C.c_double()
        rc = L.wpref_encode(0 if algo == "fast" else 1, b, len(b), arr, lens, n, C.byref(ids), C.byref(cnt),
                            C.byref(sec))
        if rc != 0:
            raise RuntimeError(L.wpref_last_error().decode())
        out = np.ctypeslib.as_array(ids, shape=(cnt.value,)).astype(np.int32) if cnt.value else np.zeros(0, np.int32)
        L.wpref_free(ids)
        return (out, sec.value) if return_seconds else out

    @classmethod
    def encode_files(cls, text_file: str, vocab_file: str, algo: str = "fast"):
        L = cls.lib()
        ids = C.POINTER(C.c_int)()
        cnt = C.c_size_t()
        sec = C.c_double()
        rc = L.wpref_encode_files(0 if algo == "fast" else 1, text_file.encode(), vocab_file.encode(),
                                  C.byref(ids), C.byref(cnt), C.byref(sec))
        if rc != 0:
            raise RuntimeError(L.wpref_last_error().decode())
        out = np.ctypeslib.as_array(ids, shape=(cnt.value,)).astype(np.int32) if cnt.value else np.zeros(0, np.int32)
        L.wpref_free(ids)
        return out, sec.value

    @classmethod
    def decode(cls, vocab_file: str, ids: Iterable[int]):
        L = cls.lib()
        a = np.asarray(list(ids), dtype=np.int32)
        buf = C.c_void_p()
        ln = C.c_size_t()
        rc = L.wpref_decode(vocab_file.encode(), a.ctypes.data_as(C.POINTER(C.c_int)), a.size, C.byref(buf),
                            C.byref(ln))
        if rc != 0:
            raise RuntimeError(L.wpref_last_error().decode())
        s = C.string_at(buf, ln.value)
        L.wpref_free(buf)
        return s.split(b"\n") if a.size and ln.value else []


def fnv1a64(ids: np.ndarray) -> int:
    """FNV-1a-64 over the little-endian bytes of an int32 id array (for logs)."""
    h = 0xCBF29CE484222325
    for byte in np.asarray(ids, dtype="<i4").tobytes():
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
