"""The CPU oracle (oracle/qr_oracle.c) against fixtures produced by the compiled reference.

tests/golden/*.npz were written by tests/golden/make_golden.py from the unmodified
reference (oracle/_ref).  Integer outputs and the decoder must be bit-identical; the
erf-based mapper functions get a tolerance because the reference calls scipy.special.erf
(noisemapper.pyx:66-67) and the oracle calls libm erf.
"""
import glob
import os

import numpy as np
import pytest

from oracle import port as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def same_bits(a, b):
    """Bit-identical float64 arrays, any NaN matching any NaN (x86 NaN sign depends on operand order)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))))


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_hamming_kats_bit_exact():
    g = load("decoder_hamming.npz")
    dec = orc.Decoder(g["vid"], g["cid"])
    assert (dec.vnum, dec.cnum, dec.ednum) == (7, 3, 12)
    for name in ("kat1", "kat2", "noconv", "zeros"):
        for mi in (0, 1, 2, 20, 50):
            ok, it, post = dec.decode(g[f"{name}_llr"], g[f"{name}_synd"], mi)
            assert ok == int(g[f"{name}_m{mi}_ok"]), (name, mi)
            assert it == int(g[f"{name}_m{mi}_it"]), (name, mi)
            assert np.array_equal(post.view(np.int64), g[f"{name}_m{mi}_post"].view(np.int64)), (name, mi)


def test_reference_known_answers():
    """Facts the reference's own test/test_decoder.py:237-266 pins (plus SURVEY 8c)."""
    g = load("decoder_hamming.npz")
    dec = orc.Decoder(g["vid"], g["cid"])
    llr = np.array([1.2, -0.8, -1.3, 1.1, -0.4, 0.5, 1.9])
    ok, it, post = dec.decode(llr, np.array([1, 1, 0], dtype=np.uint8), 20)
    assert ok == 1 and it == 0 and np.array_equal(post, llr)
    ok, it, post = dec.decode(np.array([1.05, -1.075, -1.0, 1.1, -0.4, 0.4, -0.2]),
                              np.array([1, 1, 0], dtype=np.uint8), 20)
    assert ok == 1 and it == 1
    assert list((post < 0).astype(int)) == [0, 1, 1, 0, 1, 0, 0]
    assert post[0] == float.fromhex("0x1.07c2314eb6157p+0")
    assert post[6] == float.fromhex("0x1.d034b1babb270p-6")
    mat = orc.Matrix(g["vid"], g["cid"])
    assert list(mat.eval_syndrome(np.array([0, 1, 1, 0, 1, 0, 0], dtype=np.uint8))) == [1, 1, 0]


def test_node_ops_bit_exact():
    g = load("decoder_nodes.npz")
    dec = orc.Decoder(g["vid"], g["cid"])
    c2v = np.zeros(dec.ednum)
    for c in range(dec.cnum):
        dec.process_check_node(c, g["synd"], c2v, g["v2c"])
    assert np.array_equal(c2v.view(np.int64), g["c2v"].view(np.int64))
    v2c = np.zeros(dec.ednum); post = np.zeros(dec.vnum)
    for v in range(dec.vnum):
        dec.process_var_node(v, g["llr"], g["c2v_in"], v2c, post)
    assert np.array_equal(v2c.view(np.int64), g["v2c_out"].view(np.int64))
    assert np.array_equal(post.view(np.int64), g["post"].view(np.int64))
    chk = [dec.check_lappr(g["lap"][i], g["syn"][i]) for i in range(g["lap"].shape[0])]
    assert chk == list(g["chk"])
    assert 0 < sum(chk) < len(chk)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "mapper_*.npz"))))
def test_mapper_against_reference(path):
    g = np.load(path)
    bps = int(g["bps"])
    pa = orc.PAMAlphabet(bps, float(g["step"]))
    assert np.array_equal(pa.constellation, g["constellation"])
    assert np.array_equal(pa.thresholds, g["thresholds"])
    assert np.array_equal(pa.probabilities, g["probabilities"])
    assert pa.variance == float(g["variance"])
    assert np.array_equal(pa.s_to_b, g["s_to_b"])
    nm = orc.NoiseMapper(pa, float(g["noise_var"]), g["sign_config"])
    np.testing.assert_allclose(nm.F_Y_thresholds, g["F_Y_thresholds"], rtol=0, atol=2e-16)
    np.testing.assert_allclose(nm.delta_F_Y, g["delta_F_Y"], rtol=0, atol=4e-16)
    # tables built with libc erf in the reference too: identical
    assert np.array_equal(nm.fwrd_transition_probability, g["fwrd"])
    assert np.array_equal(nm.back_transition_probability, g["back"])
    assert np.array_equal(nm.bare_llr_table, g["bare_llr_table"])
    assert np.array_equal(nm.inf_erf_table, g["inf_erf_table"])
    # integer outputs: bit exact
    idx = nm.hard_decide_index(g["y"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(pa.demap_symbols_to_bits(idx), g["bits"])
    # erf-based outputs
    np.testing.assert_allclose(nm.map_noise(g["y"], idx), g["n_hat"], rtol=0, atol=5e-15)
    np.testing.assert_allclose(nm.demap_lappr_array(g["n_hat"], g["x"]), g["lappr"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(nm.demap_lappr_array(g["n_grid"], g["j_grid"]), g["lappr_grid"],
                               rtol=1e-9, atol=1e-9)
    yh = np.array([[nm.g_inv_search(float(nv), i) for i in range(pa.order)] for nv in g["n_grid"][:8]])
    np.testing.assert_allclose(yh, g["yhat_grid"], rtol=0, atol=2e-9)
    assert np.array_equal(nm.bare_llr(g["x"]), g["bare"])
    # +-100 / +-1000 underflow every exp: log(0)-log(0) = nan in the reference too
    assert np.array_equal(orc.direct_llr(g["y"], pa, 2 * float(g["noise_var"])), g["direct"], equal_nan=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "chain_*.npz"))))
def test_chain_against_reference(path):
    """Frames chained as sims/reconciliation.pyx:127-153 does; the decoder is fed the
    REFERENCE's LLRs so its outputs must be bit-identical."""
    g = np.load(path)
    bps = int(g["bps"])
    dec = orc.Decoder(g["vid"], g["cid"]); mat = orc.Matrix(g["vid"], g["cid"])
    pa = orc.PAMAlphabet(bps, 2.0)
    seen = set()
    for si in range(len(g["snrs"])):
        nm = orc.NoiseMapper(pa, float(g[f"s{si}_noise_var"]), g["sign_config"])
        F = g[f"s{si}_x"].shape[0]
        for f in range(F):
            y = g[f"s{si}_y"][f]; x = g[f"s{si}_x"][f]
            xh = nm.hard_decide_index(y)
            assert np.array_equal(xh, g[f"s{si}_x_hat"][f])
            word = pa.demap_symbols_to_bits(xh)
            assert np.array_equal(word, g[f"s{si}_word"][f])
            assert np.array_equal(mat.eval_syndrome(word), g[f"s{si}_synd"][f])
            np.testing.assert_allclose(nm.map_noise(y, xh), g[f"s{si}_n_hat"][f], rtol=0, atol=5e-15)
            np.testing.assert_allclose(nm.demap_lappr_array(g[f"s{si}_n_hat"][f], x), g[f"s{si}_lappr"][f],
                                       rtol=1e-9, atol=1e-9)
            for mode, lk, sk in (("", "lappr", "synd"), ("hard_", "hard_lappr", "synd"),
                                 ("dir_", "dir_lappr", "dir_synd")):
                ok, it, post = dec.decode(g[f"s{si}_{lk}"][f], g[f"s{si}_{sk}"][f], int(g["maxiter"]))
                assert ok == g[f"s{si}_{mode}ok"][f] and it == g[f"s{si}_{mode}it"][f]
                assert same_bits(post, g[f"s{si}_{mode}post"][f])
                seen.add((ok, it > 0))
            assert np.array_equal(mat.eval_syndrome(pa.demap_symbols_to_bits(x)), g[f"s{si}_dir_synd"][f])
    assert len(seen) >= 1
