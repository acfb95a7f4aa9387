"""ctypes front end of oracle/qr_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The classes mirror the reference's names so parity tests read like the
reference's own usage (README.md:37-104 of the reference), but everything here
runs on the CPU in plain C and numpy.
"""
import ctypes as C

import numpy as np

from . import build

_lib = None

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_long)
_bp = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)


def _p(a, ty):
    return a.ctypes.data_as(ty)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.qro_mapper_create.restype = C.c_void_p
        L.qro_mapper_create.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _bp]
        L.qro_mapper_destroy.argtypes = [C.c_void_p]
        L.qro_mapper_tables.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]
        L.qro_map_noise.argtypes = [C.c_void_p, _dp, _lp, C.c_long, _dp]
        L.qro_g_inv_search.restype = C.c_double
        L.qro_g_inv_search.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double]
        L.qro_demap_lappr_array.argtypes = [C.c_void_p, _dp, _lp, C.c_long, C.c_int, _dp]
        L.qro_bare_llr.argtypes = [C.c_void_p, _lp, C.c_long, _dp]
        L.qro_direct_llr.argtypes = [_dp, C.c_long, _dp, C.c_int, C.c_double, _dp]
        L.qro_F_Z.argtypes = [_dp, C.c_long, C.c_double, C.c_double, _dp]
        L.qro_F_Y.argtypes = [C.c_void_p, _dp, C.c_long, _dp]
        L.qro_grid.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_long, _dp, _dp]
        L.qro_g_inv.argtypes = [C.c_void_p, _bp, _dp, _dp, C.c_long, _dp, _lp, C.c_long, _dp]
        L.qro_map_noise_sign.argtypes = [C.c_void_p, _bp, _dp, _lp, C.c_long, _dp]
        L.qro_demap_simplified.argtypes = [C.c_void_p, _bp, _dp, _dp, C.c_long, _dp, _lp, C.c_long, _dp]
        L.qro_demap_sofisticated.argtypes = [C.c_void_p, _bp, _dp, _dp, C.c_long, _dp, _lp, C.c_long, _dp]
        L.qro_information.argtypes = [C.c_void_p, _bp, _dp, _dp, C.c_long, _dp, _lp, _dp, C.c_long, _bp, _dp]
        L.qro_gray_table.argtypes = [C.c_int, _bp]
        L.qro_alphabet.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp]
        L.qro_hard_decide_index.argtypes = [_dp, C.c_int, _dp, C.c_long, _lp]
        L.qro_symbols_to_bits.argtypes = [_bp, C.c_int, _lp, C.c_long, _bp]
        L.qro_eval_syndrome.argtypes = [_lp, _lp, C.c_long, _bp, C.c_long, _bp]
        L.qro_count_errors.restype = C.c_long
        L.qro_count_errors.argtypes = [_dp, _bp, C.c_long]
        L.qro_decoder_create.restype = C.c_void_p
        L.qro_decoder_create.argtypes = [_lp, _lp, C.c_long]
        L.qro_decoder_destroy.argtypes = [C.c_void_p]
        L.qro_decoder_info.argtypes = [C.c_void_p, _lp, _lp, _lp]
        L.qro_check_lappr.restype = C.c_int
        L.qro_check_lappr.argtypes = [C.c_void_p, _dp, _bp]
        L.qro_process_check_node.argtypes = [C.c_void_p, C.c_long, _bp, _dp, _dp]
        L.qro_process_var_node.argtypes = [C.c_void_p, C.c_long, _dp, _dp, _dp, _dp]
        L.qro_decode.restype = C.c_int
        L.qro_decode.argtypes = [C.c_void_p, _dp, _bp, C.c_int, _dp, _ip]
        L.qro_decode_frames.argtypes = [C.c_void_p, _dp, _bp, C.c_long, C.c_int, _dp, _bp, _ip]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def gray_table(bps):
    out = np.zeros((1 << bps, bps), dtype=np.uint8)
    lib().qro_gray_table(bps, _p(out, _bp))
    return out


def count_errors_from_lappr(lappr, word):
    lappr = _f64(lappr); word = _u8(word)
    assert lappr.size == word.size
    return int(lib().qro_count_errors(_p(lappr, _dp), _p(word, _bp), lappr.size))


class PAMAlphabet:
    def __init__(self, bit_per_symbol, step, probabilities=None):
        self.bit_per_symbol = int(bit_per_symbol)
        self.order = 1 << self.bit_per_symbol
        self.step = float(step)
        self.constellation = np.zeros(self.order)
        self.thresholds = np.zeros(self.order + 1)
        self.probabilities = np.zeros(self.order)
        var = C.c_double(0)
        pin = None if probabilities is None else _p(_f64(probabilities), _dp)
        rc = lib().qro_alphabet(self.bit_per_symbol, self.step, pin, _p(self.constellation, _dp),
                                _p(self.thresholds, _dp), _p(self.probabilities, _dp), C.byref(var))
        if rc:
            raise ValueError("bad alphabet")
        self.variance = var.value
        self.s_to_b = gray_table(self.bit_per_symbol)

    def index_to_value(self, index):
        return self.constellation[_i64(index)]

    def demap_symbols_to_bits(self, symbol_index):
        idx = _i64(symbol_index)
        out = np.zeros(idx.size * self.bit_per_symbol, dtype=np.uint8)
        lib().qro_symbols_to_bits(_p(self.s_to_b, _bp), self.bit_per_symbol, _p(idx, _lp),
                                  idx.size, _p(out, _bp))
        return out


def F_Z(z, mu, sigma):
    z = _f64(z)
    out = np.zeros(z.size)
    lib().qro_F_Z(_p(z, _dp), z.size, float(mu), float(sigma), _p(out, _dp))
    return out


class NoiseMapper:
    def __init__(self, pa, noise_var, sign_config=None, trunkation_threshold=1e-21, n_intervals_per_step=1000):
        self.pa = pa
        self.order = pa.order
        self.bit_per_symbol = pa.bit_per_symbol
        self.noise_var = float(noise_var)
        self.noise_sigma = float(np.sqrt(noise_var))
        self.constellation = pa.constellation
        self.thresholds = pa.thresholds
        self.probabilities = pa.probabilities
        sc = np.zeros(pa.order, dtype=np.uint8) if sign_config is None else _u8(sign_config)[:pa.order].copy()
        self.sign_config = sc
        self.half_order = pa.order >> 1
        self._sign_g = self._g_signs(sc)
        # the dense grid of noisemapper.pyx:135-144, built on first use
        if trunkation_threshold > 1.0:
            self._y_low, self._y_high = pa.constellation[0] * 10, pa.constellation[-1] * 10
        else:
            tmp = np.sqrt(-2.0 * np.log(trunkation_threshold)) * float(np.sqrt(noise_var))
            self._y_high, self._y_low = pa.constellation[-1] + tmp, pa.constellation[0] - tmp
        self._n_points = int(np.ceil((self._y_high - self._y_low) * n_intervals_per_step / pa.step)) + 1
        self._grid = None
        self._h = lib().qro_mapper_create(pa.bit_per_symbol, _p(pa.constellation, _dp),
                                          _p(pa.thresholds, _dp), _p(pa.probabilities, _dp),
                                          self.noise_var, _p(sc, _bp))
        if not self._h:
            raise ValueError("bad mapper arguments")
        M, b = pa.order, pa.bit_per_symbol
        self.F_Y_thresholds = np.zeros(M + 1)
        self.delta_F_Y = np.zeros(M)
        self.fwrd_transition_probability = np.zeros((M, M))
        self.back_transition_probability = np.zeros((M, M))
        self.bare_llr_table = np.zeros((M, b))
        self.inf_erf_table = np.zeros((M, M))
        lib().qro_mapper_tables(self._h, _p(self.F_Y_thresholds, _dp), _p(self.delta_F_Y, _dp),
                                _p(self.fwrd_transition_probability, _dp),
                                _p(self.back_transition_probability, _dp),
                                _p(self.bare_llr_table, _dp), _p(self.inf_erf_table, _dp))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().qro_mapper_destroy(h)
            self._h = None

    def _g_signs(self, sc):
        """sign vector g / g_inv work with (the subclasses override it, noisemapper.pyx:775-816)"""
        return sc

    def _build_grid(self):
        if self._grid is None:
            y = np.zeros(self._n_points); F = np.zeros(self._n_points)
            lib().qro_grid(self._h, float(self._y_low), float(self._y_high), self._n_points, _p(y, _dp), _p(F, _dp))
            self._grid = (y, F)
        return self._grid

    @property
    def y_range(self):
        return self._build_grid()[0].copy()

    @property
    def F_Y_values(self):
        return self._build_grid()[1].copy()

    def F_Y(self, y):
        y = _f64(y)
        out = np.zeros(y.size)
        lib().qro_F_Y(self._h, _p(y, _dp), y.size, _p(out, _dp))
        return out

    def index_to_val(self, index):
        return self.constellation[_i64(index)]

    def demap_noise(self, n_hat, symb):
        n_hat = _f64(n_hat); symb = _i64(symb)
        if n_hat.size != symb.size:
            raise ValueError("Sizes do not match")
        gy, gF = self._build_grid()
        out = np.zeros(n_hat.size)
        lib().qro_g_inv(self._h, _p(self._sign_g, _bp), _p(gF, _dp), _p(gy, _dp), gy.size, _p(n_hat, _dp),
                        _p(symb, _lp), n_hat.size, _p(out, _dp))
        return out

    def g_inv(self, n_hat, i):
        return float(self.demap_noise([n_hat], [i])[0])

    def g(self, y, i):
        return float(self.map_noise([y], [i])[0])

    def _variant(self, fn, n, j):
        n = _f64(n); j = _i64(j)
        if n.size != j.size:
            raise ValueError("Sizes of transformed noise vector and tx symbols do not match")
        gy, gF = self._build_grid()
        out = np.zeros(n.size * self.bit_per_symbol)
        fn(self._h, _p(self._sign_g, _bp), _p(gF, _dp), _p(gy, _dp), gy.size, _p(n, _dp), _p(j, _lp), n.size,
           _p(out, _dp))
        return out

    def demap_lappr_simplified_array(self, n, j):
        return self._variant(lib().qro_demap_simplified, n, j)

    def demap_lappr_sofisticated_array(self, n, j):
        return self._variant(lib().qro_demap_sofisticated, n, j)

    def hard_decide_index(self, y):
        y = _f64(y)
        out = np.zeros(y.size, dtype=np.int64)
        lib().qro_hard_decide_index(_p(self.thresholds, _dp), self.order, _p(y, _dp), y.size, _p(out, _lp))
        return out

    def map_noise(self, y, index):
        y = _f64(y); index = _i64(index)
        if y.size != index.size:
            raise ValueError("Input vectors sizes do not match")
        out = np.zeros(y.size)
        lib().qro_map_noise_sign(self._h, _p(self._sign_g, _bp), _p(y, _dp), _p(index, _lp), y.size, _p(out, _dp))
        return out

    def g_inv_search(self, n_hat, i, y_accuracy=1e-9):
        return float(lib().qro_g_inv_search(self._h, float(n_hat), int(i), float(y_accuracy)))

    def demap_lappr_array(self, n, j, corrected=False):
        n = _f64(n); j = _i64(j)
        if n.size != j.size:
            raise ValueError("Sizes of transformed noise vector and tx symbols do not match")
        out = np.zeros(n.size * self.bit_per_symbol)
        lib().qro_demap_lappr_array(self._h, _p(n, _dp), _p(j, _lp), n.size, int(corrected), _p(out, _dp))
        return out

    def bare_llr(self, symb):
        symb = _i64(symb)
        out = np.zeros(symb.size * self.bit_per_symbol)
        lib().qro_bare_llr(self._h, _p(symb, _lp), symb.size, _p(out, _dp))
        return out


class NoiseMapperFlipSign(NoiseMapper):
    """noisemapper.pyx:775-797: g / g_inv decreasing on the lower half of the alphabet"""

    def _g_signs(self, sc):
        out = np.zeros(self.order, dtype=np.uint8)
        out[: self.half_order] = 1
        return out


class NoiseMapperAntiFlipSign(NoiseMapper):
    """noisemapper.pyx:798-816: g / g_inv decreasing on the upper half of the alphabet"""

    def _g_signs(self, sc):
        out = np.zeros(self.order, dtype=np.uint8)
        out[self.half_order:] = 1
        return out


def P_xhat(nm):
    """mutual_information.pyx:29-39"""
    return np.array([sum(nm.probabilities[j] * nm.fwrd_transition_probability[j, i] for j in range(nm.order))
                     for i in range(nm.order)])


def information_from_samples(nm, p_Xhat, x_ind, y, which=(1, 1, 1)):
    """mutual_information.pyx:241-298 for given samples -> (I_X_Xhat, I_X_Y, I_XN_Xhat)"""
    x_ind = _i64(x_ind); y = _f64(y); p_Xhat = _f64(p_Xhat); which = _u8(which)
    gy, gF = nm._build_grid()
    out = np.zeros(3)
    lib().qro_information(nm._h, _p(nm._sign_g, _bp), _p(gF, _dp), _p(gy, _dp), gy.size, _p(p_Xhat, _dp),
                          _p(x_ind, _lp), _p(y, _dp), y.size, _p(which, _bp), _p(out, _dp))
    return tuple(out)


def direct_llr(y, pa, two_variance):
    y = _f64(y)
    out = np.zeros(y.size * pa.bit_per_symbol)
    lib().qro_direct_llr(_p(y, _dp), y.size, _p(pa.constellation, _dp), pa.bit_per_symbol,
                         float(two_variance), _p(out, _dp))
    return out


class Matrix:
    def __init__(self, vid, cid):
        self.vid = _i64(vid).copy(); self.cid = _i64(cid).copy()
        if self.vid.size != self.cid.size:
            raise ValueError("Incompatible sizes for input vectors")
        self.ednum = self.vid.size
        self.cnum = int(self.cid.max()) + 1
        self.vnum = int(self.vid.max()) + 1

    def eval_syndrome(self, word):
        word = _u8(word)
        if word.size < self.vnum:
            raise IndexError("word shorter than the number of variable nodes")
        out = np.zeros(self.cnum, dtype=np.uint8)
        lib().qro_eval_syndrome(_p(self.vid, _lp), _p(self.cid, _lp), self.ednum, _p(word, _bp),
                                self.cnum, _p(out, _bp))
        return out


class Decoder:
    def __init__(self, vid, cid):
        vid = _i64(vid); cid = _i64(cid)
        if vid.size != cid.size:
            raise ValueError("Sizes don't match")
        self._h = lib().qro_decoder_create(_p(vid, _lp), _p(cid, _lp), vid.size)
        n, c, e = C.c_long(), C.c_long(), C.c_long()
        lib().qro_decoder_info(self._h, C.byref(n), C.byref(c), C.byref(e))
        self.vnum, self.cnum, self.ednum = n.value, c.value, e.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().qro_decoder_destroy(h)
            self._h = None

    def check_lappr(self, lappr, synd):
        lappr = _f64(lappr); synd = _u8(synd)
        return int(lib().qro_check_lappr(self._h, _p(lappr, _dp), _p(synd, _bp)))

    def process_check_node(self, node, synd, c2v, v2c):
        """In place on c2v (float64, contiguous), like the reference's debug entry point."""
        synd = _u8(synd)
        assert c2v.dtype == np.float64 and c2v.flags.c_contiguous
        lib().qro_process_check_node(self._h, int(node), _p(synd, _bp), _p(c2v, _dp), _p(_f64(v2c), _dp))

    def process_var_node(self, node, llr, c2v, v2c, post):
        assert v2c.dtype == np.float64 and post.dtype == np.float64
        lib().qro_process_var_node(self._h, int(node), _p(_f64(llr), _dp), _p(_f64(c2v), _dp),
                                   _p(v2c, _dp), _p(post, _dp))

    def decode(self, llr, synd, max_iterations):
        llr = _f64(llr); synd = _u8(synd)
        post = np.zeros(self.vnum)
        it = C.c_int(0)
        ok = lib().qro_decode(self._h, _p(llr, _dp), _p(synd, _bp), int(max_iterations),
                              _p(post, _dp), C.byref(it))
        return int(ok), it.value, post

    def decode_frames(self, llr, synd, max_iterations):
        llr = _f64(llr).reshape(-1, self.vnum); synd = _u8(synd).reshape(-1, self.cnum)
        F = llr.shape[0]
        post = np.zeros((F, self.vnum)); ok = np.zeros(F, dtype=np.uint8); it = np.zeros(F, dtype=np.int32)
        lib().qro_decode_frames(self._h, _p(llr, _dp), _p(synd, _bp), F, int(max_iterations),
                                _p(post, _dp), _p(ok, _bp), _p(it, _ip))
        return ok, it, post
