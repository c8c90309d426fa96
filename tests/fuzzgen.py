"""ctypes front-end of tests/cpp/fuzzgen.cpp (the reference fuzzers' input generators, same seeds / same libstdc++ streams)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "fuzzgen.cpp")
SO = os.path.join(HERE, "cpp", "libfuzzgen.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-fvisibility=hidden", "-o", SO, SRC])
    return SO


_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(build())
        L.fuzzgen_mutate.restype = C.c_int
        L.fuzzgen_mutate.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.fuzzgen_dims_open.restype = C.c_void_p
        L.fuzzgen_dims_close.argtypes = [C.c_void_p]
        L.fuzzgen_dims_next.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 3
        L.fuzzgen_dims_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.fuzzgen_dims_skip.argtypes = [C.c_void_p, C.c_uint64]
        _L = L
    return _L


def mutate(trial: int, source: np.ndarray, chans: int):
    """fuzz_test_encoder trial `trial` applied to a copy of `source` (flat uint8). Returns (buffer, family 0..5)."""
    buf = np.ascontiguousarray(source, dtype=np.uint8).reshape(-1).copy()
    fam = lib().fuzzgen_mutate(trial, buf.ctypes.data_as(C.c_void_p), buf.size, chans)
    return buf, fam


class DimSession:
    """fuzz_test_encoder2's stream of (w, h, chans, pixels): trials must be consumed in order."""

    def __init__(self):
        self.s = lib().fuzzgen_dims_open()

    def next(self, want_pixels: bool = True):
        w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().fuzzgen_dims_next(self.s, C.byref(w), C.byref(h), C.byref(c))
        w, h, c = w.value, h.value, c.value
        if not want_pixels:
            lib().fuzzgen_dims_skip(self.s, w * h)
            return w, h, c, None
        buf = np.empty(w * h * c, dtype=np.uint8)
        lib().fuzzgen_dims_fill(self.s, buf.ctypes.data_as(C.c_void_p), w * h, c)
        return w, h, c, buf

    def close(self):
        if self.s:
            lib().fuzzgen_dims_close(self.s)
            self.s = None
