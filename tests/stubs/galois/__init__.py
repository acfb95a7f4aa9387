"""Minimal stand-in for the third-party `galois` package (absent from this image), exactly as much of it as
the reference's test/test_decoder.py uses: GF2(array) with field addition (XOR), indexing, truth value of
elements, and GF2.Random(n).  Test infrastructure only; authored here (no galois source)."""
import numpy as np


class GF2(np.ndarray):
    def __new__(cls, values):
        a = np.asarray(values)
        if a.size and (a.min() < 0 or a.max() > 1):
            raise ValueError("GF(2) elements are 0 or 1")
        return a.astype(np.uint8).view(cls)

    @classmethod
    def Random(cls, shape, seed=None):
        return cls(np.random.default_rng(seed).integers(0, 2, size=shape))

    def __add__(self, other):
        return GF2(np.bitwise_xor(np.asarray(self), np.asarray(other, dtype=np.uint8)))

    __radd__ = __add__
    __sub__ = __add__

    def __mul__(self, other):
        return GF2(np.bitwise_and(np.asarray(self), np.asarray(other, dtype=np.uint8)))
