"""Softening noise mapper and LLR demapper (reference: qamreconciliation/noisemapper.pyx)."""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _abi
from ._util import device, dtype_code, stream, to_dev, to_np
from .alphabet import PAMAlphabet


def _demap_mode(mode):
    if mode is None:
        mode = os.environ.get("QAMRECON_DEMAP", "exact")
    if isinstance(mode, int):
        return mode
    table = {"exact": _abi.QR_DEMAP_EXACT, "fast": _abi.QR_DEMAP_FAST, "fast64": _abi.QR_DEMAP_FAST,
             "exact-corrected": _abi.QR_DEMAP_EXACT | _abi.QR_DEMAP_CORRECTED,
             "fast-corrected": _abi.QR_DEMAP_FAST | _abi.QR_DEMAP_CORRECTED,
             # LLRs to float precision (what the fp32 decoder consumes): no 2^-30 cell replay, MUFU exp / log
             "fast32": _abi.QR_DEMAP_FAST | _abi.QR_DEMAP_F32GRADE,
             "fast32-corrected": _abi.QR_DEMAP_FAST | _abi.QR_DEMAP_F32GRADE | _abi.QR_DEMAP_CORRECTED}
    if mode not in table:
        raise ValueError(f"unknown demap mode {mode!r}")
    return table[mode]


class NoiseMapper:
    """NoiseMapper(pa, noise_var, sign_config=None, trunkation_threshold=1e-21, n_intervals_per_step=1000)
    -- noisemapper.pyx:103-236.  Readonly attributes as in noisemapper.pxd:19-35.

    Per-frame methods return numpy arrays like the reference's memoryviews; the `*_batch` methods
    work on CUDA tensors of any leading shape.  `demap` modes: 'exact' replays the reference's 1e-9
    bisection, 'fast' solves by Newton and lands in the same cell; '+corrected' divides the k<j
    exponent by 2 sigma^2 as well (the reference does not: noisemapper.pyx:503-507)."""

    def __init__(self, pa, noise_var, sign_config=None, trunkation_threshold=1e-21, n_intervals_per_step=1000):
        if not isinstance(pa, PAMAlphabet):
            raise TypeError("pa must be a PAMAlphabet")
        noise_var = float(noise_var)
        if noise_var <= 0:
            raise ValueError(f"noise variance must be strictly positive, got {noise_var}")
        if sign_config is None:
            self.sign_config = np.zeros(pa.order, dtype=np.uint8)
        else:
            sc = np.array(np.asarray(sign_config), dtype=np.uint8, copy=True).ravel()
            if sc.size < pa.order:
                raise ValueError("Not enough data for a monotonicity sign configuration")
            self.sign_config = sc
        self.order = pa.order
        self.half_order = pa.order >> 1
        self.bit_per_symbol = pa.bit_per_symbol
        self.constellation = pa.constellation
        self.variance = pa.variance
        self.thresholds = pa.thresholds
        self.probabilities = pa.probabilities
        self.noise_var = noise_var
        self.noise_sigma = math.sqrt(noise_var)
        self._pa = pa
        # the dense F_Y grid of noisemapper.pyx:135-144 is only read by g_inv: built on first use
        if trunkation_threshold > 1.0:
            self._y_low = pa.constellation[0] * 10
            self._y_high = pa.constellation[-1] * 10
        else:
            tmp = math.sqrt(-2.0 * math.log(trunkation_threshold)) * self.noise_sigma
            self._y_high = pa.constellation[-1] + tmp
            self._y_low = pa.constellation[0] - tmp
        self._n_points = int(math.ceil((self._y_high - self._y_low) * n_intervals_per_step / pa.step)) + 1
        self._grid = None

        dev = device()
        h = C.c_void_p()
        sc = np.ascontiguousarray(self.sign_config[: pa.order])
        _abi.check(_abi.lib().qr_mapper_create(
            pa.bit_per_symbol, np.ascontiguousarray(pa.constellation).ctypes.data,
            np.ascontiguousarray(pa.thresholds).ctypes.data, np.ascontiguousarray(pa.probabilities).ctypes.data,
            noise_var, sc.ctypes.data, dev.index, C.byref(h)))
        self._h = h
        M, b = pa.order, pa.bit_per_symbol
        self.F_Y_thresholds = np.empty(M + 1)
        self.delta_F_Y = np.empty(M)
        self.fwrd_transition_probability = np.empty((M, M))
        self.back_transition_probability = np.empty((M, M))
        self.bare_llr_table = np.empty((M, b))
        self.inf_erf_table = np.empty((M, M))
        _abi.check(_abi.lib().qr_mapper_tables(
            h, self.F_Y_thresholds.ctypes.data, self.delta_F_Y.ctypes.data,
            self.fwrd_transition_probability.ctypes.data, self.back_transition_probability.ctypes.data,
            self.bare_llr_table.ctypes.data, self.inf_erf_table.ctypes.data))
        sg = self._g_signs()
        if sg is not None:
            sg = np.ascontiguousarray(sg, dtype=np.uint8)
            _abi.check(_abi.lib().qr_mapper_set_g_sign(h, sg.ctypes.data))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _abi.lib().qr_mapper_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- caller-supplied symbol indices -----------------------------------------------------------
    def check_indices(self):
        """Raise IndexError (as the reference's bounds-checked Cython does) if any symbol / region index handed to
        this mapper since the last check was outside [0, order).  Synchronises the current stream; the per-frame
        (numpy) methods call it themselves, users of the `*_batch` methods call it when they next synchronise."""
        n = C.c_int64()
        _abi.check(_abi.lib().qr_mapper_index_errors(self._h, C.byref(n), stream()))
        if n.value:
            raise IndexError(f"{n.value} symbol indices out of bounds for an alphabet of order {self.order}")

    def _np_checked(self, t):
        out = to_np(t)
        self.check_indices()
        return out

    # -- sign rule of g / g_inv (overridden by the FlipSign subclasses, noisemapper.pyx:775-816) ------
    def _g_signs(self):
        return None

    # -- lazily built grid (noisemapper.pyx:135-144, :254-261), computed and kept on the device ------
    def _build_grid(self):
        if self._grid is None:
            _abi.check(_abi.lib().qr_mapper_build_grid(self._h, float(self._y_low), float(self._y_high), self._n_points))
            y = np.empty(self._n_points); F = np.empty(self._n_points)
            _abi.check(_abi.lib().qr_mapper_grid(self._h, None, y.ctypes.data, F.ctypes.data))
            self._grid = (y, F)
        return self._grid

    @property
    def y_range(self):
        return np.array(self._build_grid()[0])

    @property
    def F_Y_values(self):
        return np.array(self._build_grid()[1])

    def F_Y_batch(self, y):
        t = to_dev(y, torch.float64)
        out = torch.empty(t.shape, dtype=torch.float64, device=t.device)
        _abi.check(_abi.lib().qr_F_Y(self._h, t.data_ptr(), t.numel(), out.data_ptr(), stream()))
        return out

    def F_Y(self, y):
        """noisemapper.pyx:264-275: (sum of the Gaussian CDFs) / order"""
        return to_np(self.F_Y_batch(to_dev(y, torch.float64).reshape(-1)))

    def demap_noise_batch(self, n_hat, symb):
        """g_inv per element: interpolation on the F_Y grid (noisemapper.pyx:391-404, :295-307, :47-63)"""
        nn = to_dev(n_hat, torch.float64); ss = to_dev(symb, torch.int64)
        if nn.numel() != ss.numel():
            raise ValueError("Sizes do not match")
        self._build_grid()
        out = torch.empty(nn.shape, dtype=torch.float64, device=nn.device)
        _abi.check(_abi.lib().qr_demap_noise(self._h, nn.data_ptr(), ss.data_ptr(), nn.numel(), out.data_ptr(), stream()))
        return out

    def demap_noise(self, n_hat, symb):
        """noisemapper.pyx:391-404"""
        return self._np_checked(self.demap_noise_batch(to_dev(n_hat, torch.float64).reshape(-1),
                                            to_dev(symb, torch.int64).reshape(-1)))

    def g_inv(self, n_hat, i):
        """noisemapper.pyx:295-307"""
        return float(self.demap_noise(np.array([float(n_hat)]), np.array([int(i)], dtype=np.int64))[0])

    def _variant_batch(self, variant, n, j):
        nn = to_dev(n, torch.float64); jj = to_dev(j, torch.int64)
        if nn.numel() != jj.numel():
            raise ValueError("Sizes of transformed noise vector and tx symbols do not match")
        self._build_grid()
        out = torch.empty(nn.shape[:-1] + (nn.shape[-1] * self.bit_per_symbol,), dtype=torch.float64, device=nn.device)
        _abi.check(_abi.lib().qr_demap_lappr_variant(self._h, variant, nn.data_ptr(), jj.data_ptr(), nn.numel(),
                                                     out.data_ptr(), stream()))
        return out

    def demap_lappr_simplified_array_batch(self, n, j):
        return self._variant_batch(1, n, j)

    def demap_lappr_sofisticated_array_batch(self, n, j):
        return self._variant_batch(2, n, j)

    def demap_lappr_simplified_array(self, n, j):
        """noisemapper.pyx:605-621"""
        return self._np_checked(self._variant_batch(1, to_dev(n, torch.float64).reshape(-1), to_dev(j, torch.int64).reshape(-1)))

    def demap_lappr_simplified(self, n, j):
        """noisemapper.pyx:563-601"""
        return self.demap_lappr_simplified_array(np.array([float(n)]), np.array([int(j)], dtype=np.int64))

    def demap_lappr_sofisticated_array(self, n, j):
        """noisemapper.pyx:751-766"""
        return self._np_checked(self._variant_batch(2, to_dev(n, torch.float64).reshape(-1), to_dev(j, torch.int64).reshape(-1)))

    def demap_lappr_sofisticated(self, n, j):
        """noisemapper.pyx:624-748"""
        return self.demap_lappr_sofisticated_array(np.array([float(n)]), np.array([int(j)], dtype=np.int64))

    # -- hot path, batched ---------------------------------------------------------------------------
    def hard_decide_index_batch(self, y_samples):
        y = to_dev(y_samples, torch.float64)
        idx = torch.empty(y.shape, dtype=torch.int64, device=y.device)
        _abi.check(_abi.lib().qr_hard_decide_index(self._h, y.data_ptr(), y.numel(), idx.data_ptr(), stream()))
        return idx

    def map_noise_batch(self, y_samples, index):
        y = to_dev(y_samples, torch.float64); idx = to_dev(index, torch.int64)
        if y.numel() != idx.numel():
            raise ValueError("Input vectors sizes do not match")
        out = torch.empty(y.shape, dtype=torch.float64, device=y.device)
        _abi.check(_abi.lib().qr_map_noise(self._h, y.data_ptr(), idx.data_ptr(), y.numel(), out.data_ptr(), stream()))
        return out

    def front_end_batch(self, y_samples, want_index=True, want_noise=True, want_bits=True):
        """hard_decide_index + map_noise + demap_symbols_to_bits in one pass over y."""
        y = to_dev(y_samples, torch.float64)
        idx = torch.empty(y.shape, dtype=torch.int64, device=y.device) if want_index else None
        nh = torch.empty(y.shape, dtype=torch.float64, device=y.device) if want_noise else None
        bits = (torch.empty(y.shape[:-1] + (y.shape[-1] * self.bit_per_symbol,), dtype=torch.uint8, device=y.device)
                if want_bits else None)
        _abi.check(_abi.lib().qr_front_end(self._h, y.data_ptr(), y.numel(),
                                           idx.data_ptr() if want_index else None,
                                           nh.data_ptr() if want_noise else None,
                                           bits.data_ptr() if want_bits else None, stream()))
        return idx, nh, bits

    def demap_lappr_array_batch(self, n, j, mode=None, alpha=1.0, out_dtype=torch.float64):
        nn = to_dev(n, torch.float64); jj = to_dev(j, torch.int64)
        if nn.numel() != jj.numel():
            raise ValueError("Sizes of transformed noise vector and tx symbols do not match")
        out = torch.empty(nn.shape[:-1] + (nn.shape[-1] * self.bit_per_symbol,), dtype=out_dtype, device=nn.device)
        _abi.check(_abi.lib().qr_demap_lappr(self._h, nn.data_ptr(), jj.data_ptr(), nn.numel(), _demap_mode(mode),
                                             float(alpha), out.data_ptr(), dtype_code(out), stream()))
        return out

    def bare_llr_batch(self, symb, out_dtype=torch.float64):
        s = to_dev(symb, torch.int64)
        out = torch.empty(s.shape[:-1] + (s.shape[-1] * self.bit_per_symbol,), dtype=out_dtype, device=s.device)
        _abi.check(_abi.lib().qr_bare_llr(self._h, s.data_ptr(), s.numel(), out.data_ptr(), dtype_code(out), stream()))
        return out

    def direct_llr_batch(self, y_samples, two_variance=None, out_dtype=torch.float64):
        """sims/reconciliation.pyx:25-72 (`y_to_lappr_grey_array`)."""
        y = to_dev(y_samples, torch.float64)
        if two_variance is None:
            two_variance = 2 * self.noise_var
        out = torch.empty(y.shape[:-1] + (y.shape[-1] * self.bit_per_symbol,), dtype=out_dtype, device=y.device)
        _abi.check(_abi.lib().qr_direct_llr(self._h, y.data_ptr(), y.numel(), float(two_variance), out.data_ptr(),
                                            dtype_code(out), stream()))
        return out

    def g_inv_search_batch(self, n_hat, region, mode=None):
        nn = to_dev(n_hat, torch.float64); rr = to_dev(region, torch.int64)
        if nn.numel() != rr.numel():
            raise ValueError("Sizes do not match")
        out = torch.empty(nn.shape, dtype=torch.float64, device=nn.device)
        _abi.check(_abi.lib().qr_g_inv_search(self._h, nn.data_ptr(), rr.data_ptr(), nn.numel(), _demap_mode(mode),
                                              out.data_ptr(), stream()))
        return out

    # -- reference API (per frame, numpy out) ----------------------------------------------------------
    def hard_decide_index(self, y_samples):
        """noisemapper.pyx:349-359"""
        return to_np(self.hard_decide_index_batch(to_dev(y_samples, torch.float64).reshape(-1)))

    def index_to_val(self, index):
        """noisemapper.pyx:362-370"""
        return np.asarray(self.constellation)[np.asarray(index, dtype=np.int64)]

    def map_noise(self, y_samples, index):
        """noisemapper.pyx:373-388"""
        y = to_dev(y_samples, torch.float64).reshape(-1); idx = to_dev(index, torch.int64).reshape(-1)
        if y.numel() != idx.numel():
            raise ValueError("Input vectors sizes do not match")
        return self._np_checked(self.map_noise_batch(y, idx))

    def g(self, y, i):
        """noisemapper.pyx:289-292"""
        return float(self.map_noise(np.array([float(y)]), np.array([int(i)], dtype=np.int64))[0])

    def g_inv_search(self, n_hat, i, y_accuracy=1e-9):
        """noisemapper.pyx:310-345 (the kernels implement the default 1e-9 accuracy)."""
        if y_accuracy != 1e-9:
            raise ValueError("only y_accuracy=1e-9 is implemented on the device")
        return float(self._np_checked(self.g_inv_search_batch(np.array([float(n_hat)]), np.array([int(i)], dtype=np.int64)))[0])

    def demap_noise_search(self, n_hat, symb, y_accuracy=1e-9):
        """noisemapper.pyx:407-419"""
        if y_accuracy != 1e-9:
            raise ValueError("only y_accuracy=1e-9 is implemented on the device")
        nn = to_dev(n_hat, torch.float64).reshape(-1); ss = to_dev(symb, torch.int64).reshape(-1)
        if nn.numel() != ss.numel():
            raise ValueError("Sizes do not match")
        return self._np_checked(self.g_inv_search_batch(nn, ss))

    def bare_llr(self, symb):
        """noisemapper.pyx:423-432"""
        return self._np_checked(self.bare_llr_batch(to_dev(symb, torch.int64).reshape(-1)))

    def demap_lappr(self, n, j):
        """noisemapper.pyx:450-540 for one (n, j) pair"""
        return self._np_checked(self.demap_lappr_array_batch(np.array([float(n)]), np.array([int(j)], dtype=np.int64)))

    def demap_lappr_array(self, n, j):
        """noisemapper.pyx:544-559"""
        nn = to_dev(n, torch.float64).reshape(-1); jj = to_dev(j, torch.int64).reshape(-1)
        if nn.numel() != jj.numel():
            raise ValueError("Sizes of transformed noise vector and tx symbols do not match")
        return self._np_checked(self.demap_lappr_array_batch(nn, jj))


class NoiseDemapper(NoiseMapper):
    """Empty subclass, as in the reference (noisemapper.pxd:89-92)."""
    pass


class NoiseMapperFlipSign(NoiseMapper):
    """noisemapper.pyx:775-797: g / g_inv decreasing on the lower half of the alphabet.  As in the reference,
    g_inv_search / demap_lappr are NOT overridden and keep the constructor's sign_config."""

    def _g_signs(self):
        sg = np.zeros(self.order, dtype=np.uint8)
        sg[: self.half_order] = 1
        return sg


class NoiseMapperAntiFlipSign(NoiseMapper):
    """noisemapper.pyx:798-816: g / g_inv decreasing on the upper half of the alphabet."""

    def _g_signs(self):
        sg = np.zeros(self.order, dtype=np.uint8)
        sg[self.half_order:] = 1
        return sg


def F_Z(z, mu, sigma):
    """module function F_Z (noisemapper.pyx:70-80)"""
    t = to_dev(z, torch.float64).reshape(-1)
    out = torch.empty(t.shape, dtype=torch.float64, device=t.device)
    _abi.check(_abi.lib().qr_F_Z(t.data_ptr(), t.numel(), float(mu), float(sigma), out.data_ptr(), stream()))
    return to_np(out)
