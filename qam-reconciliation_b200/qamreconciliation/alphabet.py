"""PAM alphabet (reference: qamreconciliation/alphabet.pyx)."""
import ctypes as C

import numpy as np
import torch

from . import _abi, bicm
from ._util import device, stream, to_dev, to_np


class Alphabet:
    pass


class PAMAlphabet(Alphabet):
    """PAMAlphabet(bit_per_symbol, step, probabilities=None) -- alphabet.pyx:35-76.

    Host attributes as in the reference (alphabet.pxd:18-32): constellation, thresholds, variance,
    order, step, bit_per_symbol, s_to_b, probabilities."""

    def __init__(self, bit_per_symbol, step, probabilities=None):
        bit_per_symbol = int(bit_per_symbol)
        if bit_per_symbol == 0:
            raise ValueError(f"Bit per symbol must be at least 1, got {bit_per_symbol}")
        if not 0 < bit_per_symbol <= 8:
            raise ValueError("Bit per symbol must be in 1..8")
        self.bit_per_symbol = bit_per_symbol
        self.order = 1 << bit_per_symbol
        self.step = float(step)
        if probabilities is None:
            self.probabilities = np.ones(self.order, dtype=np.double) / self.order
        else:
            p = np.array(probabilities, dtype=np.double, copy=True).ravel()
            if p.size != self.order:
                raise ValueError("Probability vector does not match constellation size")
            if np.any(p <= 0):
                raise ValueError("Probabilities must be positive")   # the reference forgets to raise (:53-54)
            tmp = 0.0
            for v in p:
                tmp += v
            if np.abs(tmp - 1) > 1e-9:
                raise ValueError("Probabilities do not sum to 1")
            self.probabilities = p
        self.constellation = (np.arange(self.order) - (self.order - 1) / 2) * self.step
        self.variance = 0.0
        for i in range(self.order):
            self.variance += self.probabilities[i] * np.abs(self.constellation[i]) ** 2
        self.variance = float(self.variance)
        self.thresholds = np.empty(self.order + 1, dtype=np.double)
        for i in range(1, self.order):
            self.thresholds[i] = self.constellation[i] - self.step / 2
        self.thresholds[0] = self.constellation[0] * 100
        self.thresholds[-1] = self.constellation[-1] * 100
        self.s_to_b = bicm.generate_table_s_to_b(self.bit_per_symbol)
        self._h = None

    # -- device handle (only the Gray map is used through it) --------------------------------
    def _handle(self):
        if self._h is None:
            dev = device()
            h = C.c_void_p()
            _abi.check(_abi.lib().qr_mapper_create(
                self.bit_per_symbol, self.constellation.ctypes.data, self.thresholds.ctypes.data,
                self.probabilities.ctypes.data, 1.0, None, dev.index, C.byref(h)))
            self._h = h
        return self._h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _abi.lib().qr_mapper_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- host helpers kept for API parity ----------------------------------------------------
    def random_symbols(self, N):
        """alphabet.pyx:79-83 (global numpy RNG, like the reference)."""
        return np.array(np.random.choice(self.order, size=N, p=self.probabilities), dtype=np.int64)

    def index_to_value(self, index):
        """alphabet.pyx:86-95"""
        return self.constellation[np.asarray(index, dtype=np.int64)]

    # -- hot path ----------------------------------------------------------------------------
    def demap_symbols_to_bits_batch(self, symbol_index):
        """[..., S] int64 symbol indices -> [..., S*bps] uint8 bits, on the GPU."""
        idx = to_dev(symbol_index, torch.int64)
        bits = torch.empty(idx.shape[:-1] + (idx.shape[-1] * self.bit_per_symbol,), dtype=torch.uint8,
                           device=idx.device)
        _abi.check(_abi.lib().qr_symbols_to_bits(self._handle(), idx.data_ptr(), idx.numel(), bits.data_ptr(),
                                                 stream()))
        return bits

    def demap_symbols_to_bits(self, symbol_index):
        """alphabet.pyx:98-107: bits[i*bps + k] = s_to_b[index[i], k]; returns a uint8 numpy array."""
        idx = to_dev(symbol_index, torch.int64).reshape(-1)
        out = to_np(self.demap_symbols_to_bits_batch(idx))
        n = C.c_int64()
        _abi.check(_abi.lib().qr_mapper_index_errors(self._handle(), C.byref(n), stream()))
        if n.value:     # the reference's bounds-checked table lookup raises
            raise IndexError(f"{n.value} symbol indices out of bounds for an alphabet of order {self.order}")
        return out
