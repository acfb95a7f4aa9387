"""ctypes binding of libqamrecon.so (include/qamrecon.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible, every
compute entry point raises.  The library is built in-tree by `__graft_entry__.build()`
(nvcc, sm_100a) as qam-reconciliation_b200/qamreconciliation/libqamrecon.so.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (QAMRECON_LIBRARY: another build of the SAME library, for A/B timing of two builds on one box -- a development aid)
LIB_PATH = os.environ.get("QAMRECON_LIBRARY") or os.path.join(HERE, "libqamrecon.so")

QR_OK, QR_ERR_INVALID, QR_ERR_GRAPH, QR_ERR_CUDA, QR_ERR_NOMEM = 0, 1, 2, 3, 4
QR_F32, QR_F64 = 32, 64
QR_DEMAP_EXACT, QR_DEMAP_FAST, QR_DEMAP_CORRECTED, QR_DEMAP_F32GRADE = 0, 1, 2, 4
QR_SCHED_PERSISTENT, QR_SCHED_LAUNCH, QR_SCHED_FUSED, QR_SCHED_AUTO = 0, 1, 2, 3

_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "qr_abi_version": (C.c_int, []),
    "qr_last_error": (C.c_char_p, []),
    "qr_device_count": (C.c_int, [_P(C.c_int)]),
    "qr_graph_create": (C.c_int, [_vp, _vp, _i64, C.c_int, _P(_vp)]),
    "qr_graph_create_any": (C.c_int, [_vp, _vp, _i64, C.c_int, _P(_vp)]),
    "qr_graph_fused_eligible": (C.c_int, [_vp, _P(C.c_int)]),
    "qr_graph_destroy": (None, [_vp]),
    "qr_graph_info": (C.c_int, [_vp, _P(_i64), _P(_i64), _P(_i64), _P(_i32), _P(_i32)]),
    "qr_graph_export": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "qr_eval_syndrome": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "qr_check_word": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "qr_check_lappr": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _vp, _vp]),
    "qr_count_errors": (C.c_int, [_vp, C.c_int, _vp, _i64, _i64, _i64, _vp, _vp]),
    "qr_decoder_create": (C.c_int, [_vp, C.c_int, _i64, _P(_vp)]),
    "qr_decoder_destroy": (None, [_vp]),
    "qr_decoder_set_schedule": (C.c_int, [_vp, C.c_int]),
    "qr_decode_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _i32, _vp, _vp, _vp, C.c_int, _vp]),
    "qr_decoder_last_stats": (C.c_int, [_vp, _P(_i64), _P(_i64)]),
    "qr_process_check_node": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "qr_process_var_node": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "qr_check_synd_node": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "qr_mapper_create": (C.c_int, [C.c_int, _vp, _vp, _vp, _f64, _vp, C.c_int, _P(_vp)]),
    "qr_mapper_destroy": (None, [_vp]),
    "qr_mapper_index_errors": (C.c_int, [_vp, _P(_i64), _vp]),
    "qr_mapper_tables": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qr_hard_decide_index": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "qr_symbols_to_bits": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "qr_map_noise": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "qr_front_end": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qr_demap_lappr": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _f64, _vp, C.c_int, _vp]),
    "qr_g_inv_search": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp, _vp]),
    "qr_bare_llr": (C.c_int, [_vp, _vp, _i64, _vp, C.c_int, _vp]),
    "qr_direct_llr": (C.c_int, [_vp, _vp, _i64, _f64, _vp, C.c_int, _vp]),
    "qr_mapper_set_g_sign": (C.c_int, [_vp, _vp]),
    "qr_mapper_build_grid": (C.c_int, [_vp, _f64, _f64, _i64]),
    "qr_mapper_grid": (C.c_int, [_vp, _P(_i64), _vp, _vp]),
    "qr_F_Y": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "qr_F_Z": (C.c_int, [_vp, _i64, _f64, _f64, _vp, _vp]),
    "qr_demap_noise": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "qr_demap_lappr_variant": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp, _vp]),
    "qr_information_sums": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp]),
    "qr_reconcile_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _f64, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp,
                                      C.c_int, _vp, _vp, _vp, _vp]),
    "qr_reconcile_host_compact": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _f64, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp,
                                            _vp, _vp]),
    "qr_reconcile_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _f64, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp,
                                    C.c_int, _vp, _vp, _vp]),
    "qr_host_chunk_cuts": (C.c_int, [_i64, _i64, C.c_int, _P(_i64), _i32, _P(_i32)]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Map a status code to the exception type the reference raises in the same situation."""
    if rc == QR_OK:
        return
    msg = lib().qr_last_error().decode("utf-8", "replace")
    if rc in (QR_ERR_INVALID, QR_ERR_GRAPH):
        raise ValueError(msg)
    if rc == QR_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def require_cuda():
    """Device index this process computes on; raises when there is no GPU (no CPU path exists)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("qamreconciliation (B200 build) needs a CUDA device: there is no CPU fallback")
    lib()
    return torch.cuda.current_device()
