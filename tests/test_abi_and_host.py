"""CPU-side checks: the C-ABI library loads and exports what include/qamrecon.h declares, the host
logic (alphabet, Gray table, code generator, CSV format) matches the reference fixtures, and the
product never touches the oracle."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qam-reconciliation_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qamrecon.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    lib = C.CDLL(ge.build_library())
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/qamrecon.h but not exported"
    from qamreconciliation import _abi
    assert sorted(_abi.SIGNATURES) == syms
    lib.qr_abi_version.restype = C.c_int
    assert lib.qr_abi_version() == 1


def test_graph_builds_on_host_without_a_device():
    from qamreconciliation import _abi, codes
    L = _abi.lib()
    vid, cid = codes.hamming_7_4()
    h = C.c_void_p()
    assert L.qr_graph_create(vid.ctypes.data, cid.ctypes.data, vid.size, -1, C.byref(h)) == 0
    n, c, e = C.c_int64(), C.c_int64(), C.c_int64()
    mc, mv = C.c_int32(), C.c_int32()
    assert L.qr_graph_info(h, C.byref(n), C.byref(c), C.byref(e), C.byref(mc), C.byref(mv)) == 0
    assert (n.value, c.value, e.value, mc.value, mv.value) == (7, 3, 12, 4, 3)
    order = np.zeros(3, np.int32)
    assert L.qr_graph_export(h, order.ctypes.data, None, None, None, None) == 0
    assert sorted(order) == [0, 1, 2]
    L.qr_graph_destroy(h)
    # malformed graphs -> QR_ERR_GRAPH with a message (the reference would read out of bounds)
    bad_v = np.array([0, 1, 2], dtype=np.int64); bad_c = np.array([0, 0, 1], dtype=np.int64)
    assert L.qr_graph_create(bad_v.ctypes.data, bad_c.ctypes.data, 3, -1, C.byref(h)) == _abi.QR_ERR_GRAPH
    assert b"degree" in L.qr_last_error()
    with pytest.raises(ValueError):
        _abi.check(_abi.QR_ERR_GRAPH)
    # Matrix semantics (matrix.pyx:21-38): any edge list builds; the decoder refuses such a graph
    assert L.qr_graph_create_any(bad_v.ctypes.data, bad_c.ctypes.data, 3, -1, C.byref(h)) == 0
    assert L.qr_graph_info(h, C.byref(n), C.byref(c), C.byref(e), C.byref(mc), C.byref(mv)) == 0
    assert (n.value, c.value, e.value, mc.value) == (3, 2, 3, 2)
    elig = C.c_int(7)
    assert L.qr_graph_fused_eligible(h, C.byref(elig)) == 0 and elig.value == 0
    d = C.c_void_p()
    assert L.qr_decoder_create(h, _abi.QR_F64, 32, C.byref(d)) != 0      # (no device, and not decodable)
    L.qr_graph_destroy(h)


def test_compute_classes_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import qamreconciliation as qr
    from qamreconciliation import codes
    vid, cid = codes.hamming_7_4()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qr.Decoder(vid, cid)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qr.Matrix(vid, cid)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qr.NoiseMapper(qr.PAMAlphabet(2, 2), 1.0)


def test_product_never_touches_the_oracle():
    for path in glob.glob(os.path.join(PKG, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
            text = open(path).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "libqroracle" not in text and "qr_oracle" not in text and "oracle/_ref" not in text \
                and "_ref/" not in text and "libqremu" not in text, path


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "mapper_*_snr2.npz"))))
def test_alphabet_host_attributes_match_reference(path):
    import qamreconciliation as qr
    g = np.load(path)
    pa = qr.PAMAlphabet(int(g["bps"]), float(g["step"]))
    assert np.array_equal(pa.constellation, g["constellation"])
    assert np.array_equal(pa.thresholds, g["thresholds"])
    assert np.array_equal(pa.probabilities, g["probabilities"])
    assert pa.variance == float(g["variance"])
    assert np.array_equal(pa.s_to_b, g["s_to_b"])
    assert pa.order == 1 << int(g["bps"]) and pa.bit_per_symbol == int(g["bps"])


def test_alphabet_argument_errors():
    import qamreconciliation as qr
    with pytest.raises(ValueError):
        qr.PAMAlphabet(0, 2)
    with pytest.raises(ValueError):
        qr.PAMAlphabet(2, 2, np.array([0.5, 0.5]))
    with pytest.raises(ValueError):
        qr.PAMAlphabet(1, 2, np.array([0.7, 0.7]))
    pa = qr.PAMAlphabet(1, 2, np.array([0.25, 0.75]))
    assert pa.variance == 1.0
    assert set(pa.random_symbols(50)) <= {0, 1}
    assert np.array_equal(pa.index_to_value(np.array([1, 0])), [1.0, -1.0])


def test_gray_table_and_error_table():
    from qamreconciliation import bicm
    assert bicm.generate_table_s_to_b(1).tolist() == [[0], [1]]
    assert bicm.generate_table_s_to_b(2).tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]]
    t3 = bicm.generate_table_s_to_b(3)
    assert t3.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 1, 1], [1, 1, 1], [1, 0, 1], [0, 0, 1]]
    # Gray property: neighbours differ in one bit
    for b in range(1, 7):
        t = bicm.generate_table_s_to_b(b)
        assert np.all((t[1:] ^ t[:-1]).sum(axis=1) == 1)
    n = bicm.generate_error_number_table(t3)
    assert n[0, 0] == 0 and n[0, 1] == 1 and n[0, 2] == 2 and np.array_equal(n, n.T)
    with pytest.raises(ValueError):
        bicm.generate_table_s_to_b(0)


def test_code_generator_and_csv_roundtrip(tmp_path):
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(648, 3, 6, seed=1)
    assert vid.size == 1944 and np.all(np.bincount(cid) == 6) and np.all(np.bincount(vid) == 3)
    assert np.unique(cid * 648 + vid).size == vid.size            # no double edges
    v2, c2 = codes.regular_ldpc(648, 3, 6, seed=1)
    assert np.array_equal(vid, v2) and np.array_equal(cid, c2)    # seeded
    vi, ci = codes.irregular_ldpc(300, 150, [2, 3, 8], [0.5, 0.4, 0.1], seed=4)
    assert np.bincount(ci).min() >= 2 and np.unique(ci * 300 + vi).size == vi.size
    p = tmp_path / "code.csv"
    codes.write_edge_csv(p, vid, cid)
    lines = open(p).read().splitlines()
    assert lines[0] == "eid,cid,vid" and lines[1] == "1944,324,648"   # the reference's layout (test/hamming_7-4.csv)
    rv, rc = codes.read_edge_csv(p)
    assert np.array_equal(rv, vid) and np.array_equal(rc, cid)
    hv, hc = codes.hamming_7_4()
    assert hv.size == 12 and hc.max() == 2 and hv.max() == 6


def test_host_pipeline_chunk_cuts():
    """qr_host_chunk_cuts (host logic of qr_reconcile_host, no device): the cuts cover the batch, no chunk exceeds the
    resident lanes when a quarter of the batch fits them (so no lane is ever refilled), the head ramps up by at most
    3x per piece from 64 frames and the tail ends on 64 -- only those two copies are exposed."""
    import sys
    sys.path.insert(0, PKG)
    from qamreconciliation import _abi
    L = _abi.lib()

    def cuts(frames, lanes, compact=0):
        buf = (C.c_int64 * 256)(); n = C.c_int32()
        assert L.qr_host_chunk_cuts(frames, lanes, compact, buf, 256, C.byref(n)) == 0
        return list(buf[:n.value])

    assert cuts(4096, 4096) == [0, 64, 256, 832, 1856, 2880, 3712, 4032, 4096]
    assert cuts(4096, 1024) == cuts(4096, 4096)                  # quarters of 1024 frames either way
    assert cuts(4096, 4096, compact=1) == [0, 256, 1280, 2304, 3328, 3840, 4096]
    assert cuts(100, 512) == [0, 100] and cuts(0, 512) == [0]
    for frames in (1, 31, 64, 500, 1000, 2048, 4095, 4096, 10000, 65536):
        for lanes in (32, 128, 512, 1024, 4096):
            c = cuts(frames, lanes)
            assert c[0] == 0 and c[-1] == frames and all(b > a for a, b in zip(c, c[1:])), (frames, lanes, c)
            sizes = [b - a for a, b in zip(c, c[1:])]
            quarter = max(256, ((frames + 3) // 4 + 31) // 32 * 32)
            if lanes >= quarter:
                assert max(sizes) <= lanes
            if len(sizes) >= 6 and max(sizes) >= 256:             # ramped
                assert sizes[0] == 64 and sizes[-1] == 64
                assert all(b <= 3 * a for a, b in zip(sizes[:3], sizes[1:4]))
    buf = (C.c_int64 * 2)(); n = C.c_int32()
    assert L.qr_host_chunk_cuts(4096, 1024, 0, buf, 2, C.byref(n)) != 0     # cut array too small
    assert L.qr_host_chunk_cuts(-1, 1024, 0, buf, 2, C.byref(n)) != 0
