"""Batched reverse-reconciliation pass: the body of the reference's Monte-Carlo loop
(sims/reconciliation.pyx:129-153 soft reverse, :300-308 hard reverse, :214-227 soft direct) for a
whole batch of frames at once."""
import torch

from . import _abi
from ._util import stream, to_dev

SOFT_REVERSE, HARD_REVERSE, SOFT_DIRECT = 0, 1, 2


class Reconciler:
    """Binds a Decoder, a Matrix-equivalent graph and a NoiseMapper for one SNR point."""

    def __init__(self, dec, nm, mode=SOFT_REVERSE, precision="fp32", demap="fast", alpha=1.0, lanes=None,
                 schedule=None):
        if dec.vnum % nm.bit_per_symbol:
            raise ValueError(f"codeword length {dec.vnum} is not a multiple of bits per symbol {nm.bit_per_symbol}")
        self.dec, self.nm, self.mode, self.alpha = dec, nm, mode, float(alpha)
        self.precision = precision
        self.prec_code = {"fp32": _abi.QR_F32, "fp64": _abi.QR_F64}[precision]
        self.llr_dtype = torch.float32 if precision == "fp32" else torch.float64
        from .noisemapper import _demap_mode
        self.demap_code = _demap_mode(demap)
        # the fp32 decoder consumes float LLRs: the fast demapper then runs at float grade (no 2^-30 cell replay,
        # MUFU exp / log; ~1e-6 relative, tests/test_emulation.py) -- 'fast64' keeps the double-grade fast mode
        if precision == "fp32" and (self.demap_code & _abi.QR_DEMAP_FAST) and demap != "fast64":
            self.demap_code |= _abi.QR_DEMAP_F32GRADE
        self.lanes, self.schedule = lanes, schedule
        self.N, self.C = dec.vnum, dec.cnum
        self.S = self.N // nm.bit_per_symbol

    def run_device(self, y, x, max_iterations, k_info=None, want_post=True, stagewise=False):
        """y [B, S] float64 and x [B, S] int64 already on the GPU.  Returns a dict of CUDA tensors:
        success, iters, post (or None), word, synd, bit_errors (int32 per frame over the first k_info bits).

        Default: ONE call into the library (qr_reconcile_device), intermediates in library-owned scratch.
        stagewise=True runs the stages one library call at a time (same results; also returns the LLRs)."""
        y = to_dev(y, torch.float64); x = to_dev(x, torch.int64)
        if y.dim() != 2 or y.shape[1] != self.S or x.shape != y.shape:
            raise ValueError(f"y and x must both have shape [frames, {self.S}], got {tuple(y.shape)} and {tuple(x.shape)}")
        if not stagewise:
            B = y.shape[0]
            dev = y.device
            ok = torch.empty(B, dtype=torch.uint8, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
            post = torch.empty((B, self.N), dtype=self.llr_dtype, device=dev) if (want_post or k_info is not None) else None
            word = torch.empty((B, self.N), dtype=torch.uint8, device=dev)
            synd = torch.empty((B, self.C), dtype=torch.uint8, device=dev)
            errs = torch.empty(B, dtype=torch.int32, device=dev) if k_info is not None else None
            h = self.dec._handle(self.prec_code, self.lanes, self.schedule, frames=B)
            _abi.check(_abi.lib().qr_reconcile_device(
                h, self.nm._h, self.mode, self.demap_code, self.alpha, y.data_ptr(), x.data_ptr(), B,
                int(max_iterations), int(k_info if k_info is not None else 0), ok.data_ptr(), it.data_ptr(),
                post.data_ptr() if post is not None else None,
                _abi.QR_F32 if self.llr_dtype == torch.float32 else _abi.QR_F64, word.data_ptr(), synd.data_ptr(),
                errs.data_ptr() if errs is not None else None, stream()))
            return dict(success=ok, iters=it, post=post if want_post else None, word=word, synd=synd,
                        bit_errors=errs, llr=None)
        from . import utils
        nm, dec = self.nm, self.dec
        B = y.shape[0]
        if self.mode == SOFT_DIRECT:
            word = nm._pa.demap_symbols_to_bits_batch(x)
            llr = nm.direct_llr_batch(y, out_dtype=self.llr_dtype)
        elif self.mode == HARD_REVERSE:
            _, _, word = nm.front_end_batch(y, want_index=False, want_noise=False)
            llr = nm.bare_llr_batch(x, out_dtype=self.llr_dtype)
        else:
            _, n_hat, word = nm.front_end_batch(y, want_index=False)
            llr = nm.demap_lappr_array_batch(n_hat, x, mode=self.demap_code, alpha=self.alpha, out_dtype=self.llr_dtype)
        synd = torch.empty((B, self.C), dtype=torch.uint8, device=y.device)
        _abi.check(_abi.lib().qr_eval_syndrome(dec._g.h, word.data_ptr(), synd.data_ptr(), B, stream()))
        ok, it, post = dec.decode_batch(llr, synd, max_iterations, precision=self.precision, lanes=self.lanes,
                                        schedule=self.schedule, out_dtype=self.llr_dtype, return_post=True)
        errs = utils.count_errors_batch(post, word, k=k_info) if k_info is not None else None
        return dict(success=ok, iters=it, post=post if want_post else None, word=word, synd=synd, bit_errors=errs,
                    llr=llr)

    def run_host(self, y, x, max_iterations, k_info, out):
        """Host buffers in, host buffers out, through qr_reconcile_host (copies inside the call).
        y: float64 [B, S], x: int64 [B, S] CPU tensors (pinned for asynchronous copies); `out` is a dict of
        preallocated CPU tensors: success uint8[B], iters int32[B], bit_errors int32[B], optional post [B, N]
        and word uint8[B, N]."""
        for name, t, dt in (("y", y, torch.float64), ("x", x, torch.int64)):
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous CPU tensor of dtype {dt}")
        if y.dim() != 2 or y.shape[1] != self.S or x.shape != y.shape:
            raise ValueError(f"y and x must both have shape [frames, {self.S}], got {tuple(y.shape)} and {tuple(x.shape)}")
        B = y.shape[0]
        post = out.get("post")
        word = out.get("word")
        for name, dt, shape in (("success", torch.uint8, (B,)), ("iters", torch.int32, (B,)),
                                ("bit_errors", torch.int32, (B,))):
            t = out.get(name)
            if t is None or t.is_cuda or t.dtype != dt or tuple(t.shape) != shape or not t.is_contiguous():
                raise ValueError(f"out[{name!r}] must be a contiguous CPU tensor of dtype {dt} and shape {shape}")
        if post is not None and (post.is_cuda or tuple(post.shape) != (B, self.N) or not post.is_contiguous()
                                 or post.dtype not in (torch.float32, torch.float64)):
            raise ValueError(f"out['post'] must be a contiguous float CPU tensor of shape {(B, self.N)}")
        if word is not None and (word.is_cuda or tuple(word.shape) != (B, self.N) or word.dtype != torch.uint8
                                 or not word.is_contiguous()):
            raise ValueError(f"out['word'] must be a contiguous uint8 CPU tensor of shape {(B, self.N)}")
        h = self.dec._handle(self.prec_code, self.lanes, self.schedule, frames=B)
        _abi.check(_abi.lib().qr_reconcile_host(
            h, self.nm._h, self.mode, self.demap_code, self.alpha, y.data_ptr(), x.data_ptr(), B,
            int(max_iterations), int(k_info), out["success"].data_ptr(), out["iters"].data_ptr(),
            post.data_ptr() if post is not None else None,
            (_abi.QR_F64 if post.dtype == torch.float64 else _abi.QR_F32) if post is not None else _abi.QR_F32,
            word.data_ptr() if word is not None else None, out["bit_errors"].data_ptr(), stream()))
        self.nm.check_indices()      # (the call above has synchronised: this costs one 4-byte copy)
        return out

    def run_host_compact(self, y32, x8, max_iterations, k_info, out):
        """Compact wire format through qr_reconcile_host_compact: y32 float32 [B, S], x8 uint8 [B, S] CPU tensors in;
        `out`: success uint8[B], iters int32[B], bit_errors int32[B], decisions uint8[B, ceil(N / 8)] (hard decisions
        of the final LLRs, 8 per byte)."""
        for name, t, dt in (("y32", y32, torch.float32), ("x8", x8, torch.uint8)):
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous CPU tensor of dtype {dt}")
        if y32.dim() != 2 or y32.shape[1] != self.S or x8.shape != y32.shape:
            raise ValueError(f"y32 and x8 must both have shape [frames, {self.S}]")
        B = y32.shape[0]
        dec = out["decisions"]
        if dec.is_cuda or dec.dtype != torch.uint8 or tuple(dec.shape) != (B, (self.N + 7) // 8) or not dec.is_contiguous():
            raise ValueError(f"out['decisions'] must be a contiguous uint8 CPU tensor of shape {(B, (self.N + 7) // 8)}")
        h = self.dec._handle(self.prec_code, self.lanes, self.schedule, frames=B)
        _abi.check(_abi.lib().qr_reconcile_host_compact(
            h, self.nm._h, self.mode, self.demap_code, self.alpha, y32.data_ptr(), x8.data_ptr(), B,
            int(max_iterations), int(k_info), out["success"].data_ptr(), out["iters"].data_ptr(), dec.data_ptr(),
            out["bit_errors"].data_ptr(), stream()))
        self.nm.check_indices()
        return out
