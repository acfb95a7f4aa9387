"""Parity-check matrix as an edge list (reference: qamreconciliation/matrix.pyx)."""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._util import device, stream, to_dev, to_np


class _Graph:
    """Owns a qr_graph handle (CSR/CSC tables on the device)."""

    def __init__(self, vid, cid, any_graph=False):
        vid = np.array(np.asarray(vid), dtype=np.int64, copy=True, order="C").ravel()
        cid = np.array(np.asarray(cid), dtype=np.int64, copy=True, order="C").ravel()
        if vid.shape[0] != cid.shape[0]:
            raise ValueError("Sizes don't match")
        dev = device()
        h = C.c_void_p()
        create = _abi.lib().qr_graph_create_any if any_graph else _abi.lib().qr_graph_create
        _abi.check(create(vid.ctypes.data, cid.ctypes.data, vid.size, dev.index, C.byref(h)))
        self.h = h
        self.device = dev
        n, c, e = C.c_int64(), C.c_int64(), C.c_int64()
        mc, mv = C.c_int32(), C.c_int32()
        _abi.check(_abi.lib().qr_graph_info(h, C.byref(n), C.byref(c), C.byref(e), C.byref(mc), C.byref(mv)))
        self.vnum, self.cnum, self.ednum = n.value, c.value, e.value
        self.max_check_degree, self.max_var_degree = mc.value, mv.value

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            try:
                _abi.lib().qr_graph_destroy(h)
            except Exception:
                pass
            self.h = None


class Matrix:
    """Matrix(vnode_array, cnode_array) -- matrix.pyx:21-38; readonly vnum, cnum, ednum (matrix.pxd:24-27)."""

    def __init__(self, vnode_array, cnode_array):
        if np.asarray(vnode_array).shape[0] != np.asarray(cnode_array).shape[0]:
            raise ValueError("Incompatible sizes for input vectors")
        self._g = _Graph(vnode_array, cnode_array, any_graph=True)   # the reference's Matrix takes any edge list
        self.vnum, self.cnum, self.ednum = self._g.vnum, self._g.cnum, self._g.ednum

    def eval_syndrome_batch(self, words):
        """[B, N] uint8 words -> [B, C] uint8 syndromes on the GPU (matrix.pyx:55-60 per row)."""
        w = to_dev(words, torch.uint8)
        if w.dim() != 2 or w.shape[1] != self.vnum:
            raise ValueError(f"words must have shape [frames, {self.vnum}]")
        synd = torch.empty((w.shape[0], self.cnum), dtype=torch.uint8, device=w.device)
        _abi.check(_abi.lib().qr_eval_syndrome(self._g.h, w.data_ptr(), synd.data_ptr(), w.shape[0], stream()))
        return synd

    def eval_syndrome(self, word):
        """matrix.pyx:55-60; a word shorter than the number of variable nodes raises IndexError as the
        bounds-checked reference does, extra trailing entries are ignored."""
        w = to_dev(word, torch.uint8).reshape(-1)
        if w.numel() < self.vnum:
            raise IndexError("Out of bounds on buffer access (axis 0)")
        return to_np(self.eval_syndrome_batch(w[: self.vnum].reshape(1, -1))[0])
