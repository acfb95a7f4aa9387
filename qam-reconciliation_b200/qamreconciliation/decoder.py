"""Syndrome-based sum-product LDPC decoder (reference: qamreconciliation/decoder.pyx)."""
import ctypes as C
import os

import numpy as np
import torch

from . import _abi
from ._util import device, dtype_code, env_precision, stream, to_dev, to_np
from .matrix import _Graph


def _writeback(dst, src_tensor):
    """In-place semantics of the reference's debug entry points: results land in the caller's array."""
    if isinstance(dst, torch.Tensor):
        dst.copy_(src_tensor.to(dst.device, dst.dtype))
    else:
        np.asarray(dst)[...] = to_np(src_tensor)


class Decoder:
    """Decoder(e_to_v, e_to_c) -- variable ids first, check ids second (decoder.pyx:93).

    Per-frame `decode` mirrors the reference (fp64 by default; QAMRECON_PRECISION=fp32 selects the
    fast mode); `decode_batch` is the batched-frames entry point on CUDA tensors.  Unlike the
    reference, sizes are validated and every check must have degree >= 2."""

    def __init__(self, e_to_v, e_to_c):
        if np.asarray(e_to_v).size != np.asarray(e_to_c).size:
            raise ValueError("Sizes don't match")
        self._g = _Graph(e_to_v, e_to_c)
        self._dec = {}
        self._auto_cap = {}

    # -- properties (decoder.pyx:157-172) -----------------------------------------------------
    @property
    def cnum(self):
        """ Number of check nodes """
        return self._g.cnum

    @property
    def vnum(self):
        """ Number of variable nodes """
        return self._g.vnum

    @property
    def ednum(self):
        """ Number of edges """
        return self._g.ednum

    @property
    def fused_eligible(self):
        """True when the fused flooding-iteration schedule (QR_SCHED_FUSED) can run this graph."""
        e = C.c_int()
        _abi.check(_abi.lib().qr_graph_fused_eligible(self._g.h, C.byref(e)))
        return bool(e.value)

    def __del__(self):
        for h in getattr(self, "_dec", {}).values():
            try:
                _abi.lib().qr_decoder_destroy(h)
            except Exception:
                pass
        self._dec = {}

    # -- decoder handles ------------------------------------------------------------------------
    def _auto_lanes(self, precision, frames):
        """Resident frames when the caller names none: every frame of the batch its own lane, up to 4096 and a quarter
        of the free device memory, in powers of two from 512.  More lanes than frames cost nothing (the library uses
        min(lanes, frames)); fewer mean refill generations, which measured slower than one lane per frame at every
        operating point of config 2."""
        have = self._auto_cap.get(precision, 0)
        if frames is None:
            return have or 512
        want = 512
        while want < min(int(frames), 4096):
            want *= 2
        if want > have:
            w = 8 if precision == _abi.QR_F64 else 4
            per_lane = (2 * self.ednum + 2 * self.vnum) * w + self.cnum
            try:
                free = torch.cuda.mem_get_info()[0]
            except Exception:
                free = 0
            while want > 512 and want * per_lane > free // 4:
                want //= 2
        return max(want, have)

    def _handle(self, precision, lanes=None, schedule=None, frames=None):
        """Decoder handle for (precision, lanes).  lanes None: QAMRECON_LANES, else sized from `frames` (see
        _auto_lanes); that workspace only ever grows."""
        if lanes is None:
            lanes = int(os.environ.get("QAMRECON_LANES", "0"))
        if schedule is None:
            schedule = int(os.environ.get("QAMRECON_SCHEDULE", str(_abi.QR_SCHED_AUTO)))
        key = (precision, lanes)
        create = lanes
        if lanes == 0:
            create = self._auto_lanes(precision, frames)
            if key in self._dec and create > self._auto_cap.get(precision, 0):
                _abi.lib().qr_decoder_destroy(self._dec.pop(key))
        if key not in self._dec:
            h = C.c_void_p()
            _abi.check(_abi.lib().qr_decoder_create(self._g.h, precision, create, C.byref(h)))
            self._dec[key] = h
            if lanes == 0:
                self._auto_cap[precision] = create
        _abi.check(_abi.lib().qr_decoder_set_schedule(self._dec[key], schedule))
        return self._dec[key]

    # -- batched entry point ----------------------------------------------------------------------
    def decode_batch(self, lappr_data, synd, max_iterations, precision=None, out_dtype=None, lanes=None,
                     schedule=None, return_post=True):
        """Decode independent frames: lappr_data [B, N] (float32/float64), synd [B, C] (uint8).

        Returns (success uint8[B], iterations int32[B], final_lappr [B, N] or None) as CUDA tensors,
        frame by frame what `decode` returns.  precision: 'fp64' (parity with the reference) or 'fp32'."""
        if precision is None:
            prec = env_precision()
        else:
            prec = {"fp64": _abi.QR_F64, "fp32": _abi.QR_F32, 64: _abi.QR_F64, 32: _abi.QR_F32}[precision]
        if isinstance(lappr_data, torch.Tensor) and lappr_data.dtype in (torch.float32, torch.float64):
            llr = to_dev(lappr_data, lappr_data.dtype)
        else:
            llr = to_dev(lappr_data, torch.float64)
        sy = to_dev(synd, torch.uint8)
        if llr.dim() != 2 or llr.shape[1] != self.vnum:
            raise ValueError(f"lappr_data must have shape [frames, {self.vnum}]")
        if sy.dim() != 2 or sy.shape[1] != self.cnum or sy.shape[0] != llr.shape[0]:
            raise ValueError(f"synd must have shape [frames, {self.cnum}]")
        max_iterations = int(max_iterations)
        if max_iterations < 0:
            raise ValueError("max_iterations must be >= 0")
        B = llr.shape[0]
        if out_dtype is None:
            out_dtype = torch.float64 if prec == _abi.QR_F64 else torch.float32
        success = torch.empty(B, dtype=torch.uint8, device=llr.device)
        iters = torch.empty(B, dtype=torch.int32, device=llr.device)
        post = torch.empty((B, self.vnum), dtype=out_dtype, device=llr.device) if return_post else None
        h = self._handle(prec, lanes, schedule, frames=B)
        _abi.check(_abi.lib().qr_decode_batch(
            h, llr.data_ptr(), dtype_code(llr), sy.data_ptr(), B, max_iterations, success.data_ptr(),
            iters.data_ptr(), post.data_ptr() if post is not None else None,
            dtype_code(post) if post is not None else _abi.QR_F64, stream()))
        return success, iters, post

    def last_stats(self, precision="fp32", lanes=None):
        """(flooding iterations summed over the frames of the last batch, schedule steps)."""
        prec = {"fp64": _abi.QR_F64, "fp32": _abi.QR_F32}[precision]
        it, st = C.c_int64(), C.c_int64()
        _abi.check(_abi.lib().qr_decoder_last_stats(self._handle(prec, lanes), C.byref(it), C.byref(st)))
        return it.value, st.value

    # -- reference API ------------------------------------------------------------------------------
    def decode(self, lappr_data, synd, max_iterations):
        """decoder.pyx:441-455: returns (success, iterations, final_lappr float64[N])."""
        llr = to_dev(lappr_data, torch.float64).reshape(-1)
        sy = to_dev(synd, torch.uint8).reshape(-1)
        if llr.numel() != self.vnum:
            raise ValueError("Size of lappr does not match number of vnodes")
        if sy.numel() != self.cnum:
            raise ValueError("Size of synd does not match number of cnodes")
        ok, it, post = self.decode_batch(llr.reshape(1, -1), sy.reshape(1, -1), max_iterations,
                                         out_dtype=torch.float64)
        return int(ok[0].item()), int(it[0].item()), to_np(post[0])

    def check_synd_node(self, check_node_index, word, synd):
        """decoder.pyx:190-217"""
        w = to_dev(word, torch.uint8).reshape(-1); s = to_dev(synd, torch.uint8).reshape(-1)
        if w.numel() != self.vnum:
            raise ValueError("Size of word does not match number of vnodes")
        if s.numel() != self.cnum:
            raise ValueError("Size of synd does not match number of cnodes")
        ok = torch.zeros(1, dtype=torch.uint8, device=w.device)
        _abi.check(_abi.lib().qr_check_synd_node(self._g.h, int(check_node_index), w.data_ptr(), s.data_ptr(),
                                                 ok.data_ptr(), stream()))
        return int(ok.item())

    def check_word_batch(self, words, synd):
        w = to_dev(words, torch.uint8); s = to_dev(synd, torch.uint8)
        if w.dim() != 2 or w.shape[1] != self.vnum or s.shape != (w.shape[0], self.cnum):
            raise ValueError("bad shapes")
        ok = torch.empty(w.shape[0], dtype=torch.uint8, device=w.device)
        _abi.check(_abi.lib().qr_check_word(self._g.h, w.data_ptr(), s.data_ptr(), w.shape[0], ok.data_ptr(), stream()))
        return ok

    def check_word(self, word, synd):
        """decoder.pyx:220-232"""
        return int(self.check_word_batch(to_dev(word, torch.uint8).reshape(1, -1),
                                         to_dev(synd, torch.uint8).reshape(1, -1))[0].item())

    def check_lappr_batch(self, lappr, synd):
        if isinstance(lappr, torch.Tensor) and lappr.dtype == torch.float32:
            l = to_dev(lappr, torch.float32)
        else:
            l = to_dev(lappr, torch.float64)
        s = to_dev(synd, torch.uint8)
        if l.dim() != 2 or l.shape[1] != self.vnum or s.shape != (l.shape[0], self.cnum):
            raise ValueError("bad shapes")
        ok = torch.empty(l.shape[0], dtype=torch.uint8, device=l.device)
        _abi.check(_abi.lib().qr_check_lappr(self._g.h, l.data_ptr(), dtype_code(l), s.data_ptr(), l.shape[0],
                                             ok.data_ptr(), stream()))
        return ok

    def check_lappr(self, lappr, synd):
        """decoder.pyx:260-281"""
        l = to_dev(lappr, torch.float64).reshape(-1); s = to_dev(synd, torch.uint8).reshape(-1)
        if l.numel() != self.vnum:
            raise ValueError("Size of lappr does not match number of vnodes")
        if s.numel() != self.cnum:
            raise ValueError("Size of synd does not match number of cnodes")
        return int(self.check_lappr_batch(l.reshape(1, -1), s.reshape(1, -1))[0].item())

    def process_var_node(self, node_index, lappr_data, check_to_var, var_to_check, updated_lappr):
        """decoder.pyx:301-319 -- var_to_check and updated_lappr are updated in place."""
        llr = to_dev(lappr_data, torch.float64); c2v = to_dev(check_to_var, torch.float64)
        v2c = to_dev(var_to_check, torch.float64); post = to_dev(updated_lappr, torch.float64)
        if llr.numel() != self.vnum or post.numel() != self.vnum or c2v.numel() != self.ednum or v2c.numel() != self.ednum:
            raise ValueError("array sizes do not match the graph")
        _abi.check(_abi.lib().qr_process_var_node(self._g.h, int(node_index), llr.data_ptr(), c2v.data_ptr(),
                                                  v2c.data_ptr(), post.data_ptr(), stream()))
        _writeback(var_to_check, v2c)
        _writeback(updated_lappr, post)

    def process_check_node(self, node_index, synd, check_to_var, var_to_check):
        """decoder.pyx:372-388 -- check_to_var is updated in place; returns 0 like the reference."""
        s = to_dev(synd, torch.uint8); c2v = to_dev(check_to_var, torch.float64)
        v2c = to_dev(var_to_check, torch.float64)
        if s.numel() != self.cnum or c2v.numel() != self.ednum or v2c.numel() != self.ednum:
            raise ValueError("array sizes do not match the graph")
        _abi.check(_abi.lib().qr_process_check_node(self._g.h, int(node_index), s.data_ptr(), c2v.data_ptr(),
                                                    v2c.data_ptr(), stream()))
        _writeback(check_to_var, c2v)
        return 0
