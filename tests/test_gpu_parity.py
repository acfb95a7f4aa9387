"""GPU parity tests: the CUDA path (through the Python mirror of the reference API, i.e. through the
C ABI of libqamrecon.so) against the CPU oracle and the fixtures generated from the compiled
reference.  Bars: integer / byte / index outputs bit-exact; fp64 decoder: success flags and iteration
counts equal on >= 99.9 % of frames and final LLRs within 1e-9 relative; erf-based mapper outputs
within the tolerances written next to each assert (CUDA erf/exp/log differ from libm in the last ulp).
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def qr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation
    return qamreconciliation


@pytest.fixture(scope="module")
def orc():
    from oracle import port
    return port


def rel_err(a, b):
    """max |a-b| / max(|b|, 1) elementwise, NaN == NaN"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1.0))
    return float(np.max(d)) if d.size else 0.0


# ------------------------------------------------------------------ reference's own decoder tests
class GF2(np.ndarray):
    """Stand-in for galois.GF2 as the reference's tests use it (uint8 array, + is XOR)."""
    def __new__(cls, data):
        return np.asarray(data, dtype=np.uint8).view(cls)

    def __add__(self, other):
        return np.bitwise_xor(np.asarray(self), np.asarray(other)).view(GF2)

    @classmethod
    def Random(cls, n):
        return cls(np.random.randint(0, 2, size=n))


def test_reference_construction_cases(qr):
    """test/test_decoder.py:8-128 of the reference"""
    uut = qr.Decoder(np.array([0, 1, 1, 2]), np.array([0, 0, 1, 1]))
    assert (uut.cnum, uut.vnum, uut.ednum) == (2, 3, 4)
    synd0, synd1 = GF2([1, 1]), GF2([0, 1])
    word0 = GF2([[1, 0, 1], [0, 1, 0]]); word1 = GF2([[0, 0, 1], [1, 1, 0]])
    for w in (word0[0], word0[1]):
        assert uut.check_synd_node(0, w, synd0) and uut.check_synd_node(1, w, synd0)
        assert not uut.check_synd_node(0, w, synd1) and uut.check_synd_node(1, w, synd1)
        assert uut.check_word(w, synd0) and not uut.check_word(w, synd1)
    for w in (word1[0], word1[1]):
        assert uut.check_synd_node(0, w, synd1) and uut.check_synd_node(1, w, synd1)
        assert not uut.check_synd_node(0, w, synd0) and uut.check_synd_node(1, w, synd0)
        assert uut.check_word(w, synd1) and not uut.check_word(w, synd0)
    l0 = np.array([-3.4, 0.8, -0.1]); l1 = np.array([-0.77, -0.8, 0.98])
    assert uut.check_lappr(l0, synd0) and not uut.check_lappr(l0, synd1)
    assert uut.check_lappr(l1, synd1) and not uut.check_lappr(l1, synd0)
    with pytest.raises(ValueError):
        uut.check_lappr(l0[:2], synd0)
    with pytest.raises(ValueError):
        uut.check_synd_node(0, word0[0], GF2([1]))


def test_reference_processing_cases(qr):
    """test/test_decoder.py:132-220 of the reference: in-place single-node updates"""
    uut = qr.Decoder(np.array([0, 1, 3, 1, 2, 1, 3, 4]), np.array([0, 0, 0, 1, 1, 2, 2, 2]))
    rng = np.random.default_rng(3)
    c2v, v2c = rng.normal(size=uut.ednum), rng.normal(size=uut.ednum)
    llr = rng.normal(size=uut.vnum); post = np.empty_like(llr)
    uut.process_var_node(1, llr, c2v, v2c, post)
    assert v2c[1] == pytest.approx(c2v[3] + c2v[5] + llr[1], rel=1e-6)
    assert v2c[3] == pytest.approx(c2v[1] + c2v[5] + llr[1], rel=1e-6)
    assert v2c[5] == pytest.approx(c2v[1] + c2v[3] + llr[1], rel=1e-6)
    assert post[1] == pytest.approx(c2v[1] + c2v[3] + c2v[5] + llr[1], rel=1e-6)
    uut.process_var_node(2, llr, c2v, v2c, post)
    assert v2c[4] == pytest.approx(llr[2], rel=1e-6) and post[2] == pytest.approx(c2v[4] + llr[2], rel=1e-6)
    uut.process_var_node(3, llr, c2v, v2c, post)
    assert v2c[2] == pytest.approx(c2v[6] + llr[3], rel=1e-6) and v2c[6] == pytest.approx(c2v[2] + llr[3], rel=1e-6)
    s = GF2([1, 0, 1])
    for s in (GF2([1, 0, 1]), GF2([0, 1, 0])):
        uut.process_check_node(1, s, c2v, v2c)
        pre = -2 if s[1] else 2
        assert c2v[3] == pytest.approx(pre * v2c[4] / 2, rel=1e-6) and c2v[4] == pytest.approx(pre * v2c[3] / 2, rel=1e-6)
        uut.process_check_node(2, s, c2v, v2c)
        pre = -2 if s[2] else 2
        t = np.tanh(v2c / 2)
        assert c2v[5] == pytest.approx(pre * np.arctanh(t[6] * t[7]), rel=1e-6)
        assert c2v[6] == pytest.approx(pre * np.arctanh(t[5] * t[7]), rel=1e-6)
        assert c2v[7] == pytest.approx(pre * np.arctanh(t[6] * t[5]), rel=1e-6)


def test_reference_decoding_cases(qr):
    """test/test_decoder.py:225-266 of the reference (Hamming(7,4) fixture) + the probed exact values"""
    from qamreconciliation import codes
    vid, cid = codes.hamming_7_4()
    uut = qr.Decoder(vid, cid)
    lappr = np.array([1.2, -0.8, -1.3, 1.1, -0.4, 0.5, 1.9])
    ok, it, out = uut.decode(lappr, GF2([1, 1, 0]), 20)
    assert ok and it == 0 and (out != lappr).sum() == 0
    ok, it, out = uut.decode(np.array([1.05, -1.075, -1.0, 1.1, -0.4, 0.4, -0.2]), GF2([1, 1, 0]), 20)
    assert ok and it <= 20
    assert (GF2((np.array(out) < 0).astype(int)) + GF2([0, 1, 1, 0, 1, 0, 0])).sum() == 0
    assert it == 1
    want = [1.0303068940321134, -1.0553068940321133, -0.9922344254468639, 1.0617311602831085,
            -0.3337636764480326, 0.3328109515380624, 0.02833287578953264]
    assert rel_err(out, want) <= 1e-12
    assert list(qr.Matrix(vid, cid).eval_syndrome(np.array([0, 1, 1, 0, 1, 0, 0], dtype=np.uint8))) == [1, 1, 0]


# ------------------------------------------------------------------ fixtures from the compiled reference
def test_hamming_fixture_all_iteration_limits(qr):
    g = np.load(os.path.join(GOLDEN, "decoder_hamming.npz"))
    dec = qr.Decoder(g["vid"], g["cid"])
    for name in ("kat1", "kat2", "noconv", "zeros"):
        for mi in (0, 1, 2, 20, 50):
            ok, it, post = dec.decode(g[f"{name}_llr"], g[f"{name}_synd"], mi)
            assert (ok, it) == (int(g[f"{name}_m{mi}_ok"]), int(g[f"{name}_m{mi}_it"])), (name, mi)
            assert rel_err(post, g[f"{name}_m{mi}_post"]) <= 1e-9, (name, mi)
            if it == 0:   # untouched input comes back bit for bit (incl. the -0.0 -> +0.0 quirk at maxiter 0)
                assert np.array_equal(post.view(np.int64), g[f"{name}_m{mi}_post"].view(np.int64))


def test_node_ops_fixture(qr):
    g = np.load(os.path.join(GOLDEN, "decoder_nodes.npz"))
    dec = qr.Decoder(g["vid"], g["cid"])
    c2v = np.zeros(dec.ednum)
    for c in range(dec.cnum):
        dec.process_check_node(c, g["synd"], c2v, g["v2c"])
    assert rel_err(c2v, g["c2v"]) <= 1e-12
    v2c = np.zeros(dec.ednum); post = np.zeros(dec.vnum)
    for v in range(dec.vnum):
        dec.process_var_node(v, g["llr"], g["c2v_in"], v2c, post)
    assert np.array_equal(v2c, g["v2c_out"]) and np.array_equal(post, g["post"])   # adds only: exact
    chk = dec.check_lappr_batch(g["lap"], g["syn"]).cpu().numpy()
    assert list(chk) == list(g["chk"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "mapper_*.npz"))))
def test_mapper_fixture(qr, path):
    g = np.load(path)
    bps = int(g["bps"])
    pa = qr.PAMAlphabet(bps, float(g["step"]))
    nm = qr.NoiseMapper(pa, float(g["noise_var"]), g["sign_config"])
    # tables: device erf vs scipy/libm erf
    np.testing.assert_allclose(nm.F_Y_thresholds, g["F_Y_thresholds"], rtol=0, atol=4e-16)
    np.testing.assert_allclose(nm.delta_F_Y, g["delta_F_Y"], rtol=0, atol=8e-16)
    # differences of two erf values near +-1: absolute error of a few ulp of 1, whatever the entry's size
    np.testing.assert_allclose(nm.fwrd_transition_probability, g["fwrd"], rtol=1e-12, atol=3e-16)
    np.testing.assert_allclose(nm.back_transition_probability, g["back"], rtol=1e-12, atol=3e-16)
    np.testing.assert_allclose(nm.inf_erf_table, g["inf_erf_table"], rtol=0, atol=4e-16)
    fin = np.isfinite(g["bare_llr_table"]) & (np.abs(g["bare_llr_table"]) < 1e299)
    # log of ratios of those entries: tail entries of ~1e-8 carry ~1e-8 relative error
    np.testing.assert_allclose(nm.bare_llr_table[fin], g["bare_llr_table"][fin], rtol=1e-9, atol=1e-7)
    assert np.array_equal(nm.bare_llr_table[~fin], g["bare_llr_table"][~fin])
    # integers: bit exact
    idx = nm.hard_decide_index(g["y"])
    assert idx.dtype == np.int64 and np.array_equal(idx, g["idx"])
    bits = pa.demap_symbols_to_bits(idx)
    assert bits.dtype == np.uint8 and np.array_equal(bits, g["bits"])
    # softening metric: |n - n_ref| <= 1e-14 (ratio of two erf sums)
    np.testing.assert_allclose(nm.map_noise(g["y"], idx), g["n_hat"], rtol=0, atol=1e-14)
    # LLRs, exact-bisection mode: 1e-9 relative (+1e-9 absolute near 0)
    # n within 1e-9 of 0 or 1 puts some reconstructed samples > 6 sigma into a tail, where F_Y is flat
    # at the resolution of a double (n exactly 0 or 1: F_Y saturated).  There the bisection converges
    # to wherever erf's ROUNDING flips the comparison -- a property of the erf implementation (CUDA's
    # and libm's differ by up to ~3e-3 in y) -- so those entries only get a 2 % tolerance.  Such
    # samples have probability < 1e-8 per symbol.
    for nk, jk, lk in (("n_hat", "x", "lappr"), ("n_grid", "j_grid", "lappr_grid")):
        sat = np.repeat((g[nk] < 1e-9) | (g[nk] > 1 - 1e-9), bps)
        got = nm.demap_lappr_array(g[nk], g[jk])
        np.testing.assert_allclose(got[~sat], g[lk][~sat], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(got[sat], g[lk][sat], rtol=2e-2, atol=1e-9)
        fast = nm.demap_lappr_array_batch(g[nk], g[jk], mode="fast").cpu().numpy()
        np.testing.assert_allclose(fast[~sat], g[lk][~sat], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(fast[sat], g[lk][sat], rtol=2e-2, atol=1e-9)
        # fp32-grade fast mode (QR_DEMAP_FAST | QR_DEMAP_F32GRADE, what the fp32 decoder is fed): 1e-5 relative
        # + 1e-6 absolute against the compiled reference, as float32 and as float64 output
        for dt in (torch.float32, torch.float64):
            f32g = nm.demap_lappr_array_batch(g[nk], g[jk], mode="fast32", out_dtype=dt).cpu().numpy().astype(np.float64)
            ok = np.isfinite(g[lk]) & (np.abs(g[lk]) < 1e30)
            np.testing.assert_allclose(f32g[~sat & ok], g[lk][~sat & ok], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(f32g[sat & ok], g[lk][sat & ok], rtol=2e-2, atol=1e-6)
    M = pa.order
    yh = nm.g_inv_search_batch(np.repeat(g["n_grid"][:8], M), np.tile(np.arange(M), 8)).cpu().numpy().reshape(8, M)
    inner = (g["n_grid"][:8] > 1e-9) & (g["n_grid"][:8] < 1 - 1e-9)
    np.testing.assert_allclose(yh[inner], g["yhat_grid"][inner], rtol=0, atol=2e-9)
    np.testing.assert_allclose(yh[~inner], g["yhat_grid"][~inner], rtol=0, atol=0.2)   # erf rounding boundary
    assert nm.g_inv_search(float(g["n_grid"][3]), 0) == pytest.approx(g["yhat_grid"][3, 0], abs=2e-9)
    bare = nm.bare_llr(g["x"])
    fin = np.abs(g["bare"]) < 1e299
    np.testing.assert_allclose(bare[fin], g["bare"][fin], rtol=1e-9, atol=1e-7)
    assert np.array_equal(bare[~fin], g["bare"][~fin])
    np.testing.assert_allclose(nm.direct_llr_batch(g["y"]).cpu().numpy(), g["direct"], rtol=1e-12, atol=1e-12,
                               equal_nan=True)
    with pytest.raises(ValueError):
        nm.map_noise(g["y"], idx[:-1])
    with pytest.raises(ValueError):
        nm.demap_lappr_array(g["n_hat"], g["x"][:-1])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "chain_*.npz"))))
@pytest.mark.parametrize("schedule", [0, 1, 3])   # 3 = fused where the graph allows it, else persistent
def test_chain_fixture(qr, path, schedule):
    """Whole frames as sims/reconciliation.pyx:127-153 chains them; decoder fed the reference's LLRs."""
    g = np.load(path)
    bps = int(g["bps"])
    dec = qr.Decoder(g["vid"], g["cid"]); mat = qr.Matrix(g["vid"], g["cid"])
    pa = qr.PAMAlphabet(bps, 2.0)
    for si in range(len(g["snrs"])):
        nm = qr.NoiseMapper(pa, float(g[f"s{si}_noise_var"]), g["sign_config"])
        y = g[f"s{si}_y"]; x = g[f"s{si}_x"]
        idx, nh, word = nm.front_end_batch(y)
        assert np.array_equal(idx.cpu().numpy(), g[f"s{si}_x_hat"])
        assert np.array_equal(word.cpu().numpy(), g[f"s{si}_word"])
        assert np.array_equal(mat.eval_syndrome_batch(word).cpu().numpy(), g[f"s{si}_synd"])
        np.testing.assert_allclose(nh.cpu().numpy(), g[f"s{si}_n_hat"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(nm.demap_lappr_array_batch(nh, x).cpu().numpy(), g[f"s{si}_lappr"], rtol=1e-9, atol=1e-9)
        dw = pa.demap_symbols_to_bits_batch(x)
        assert np.array_equal(dw.cpu().numpy(), g[f"s{si}_dir_word"])
        assert np.array_equal(mat.eval_syndrome_batch(dw).cpu().numpy(), g[f"s{si}_dir_synd"])
        for mode, lk, sk in (("", "lappr", "synd"), ("hard_", "hard_lappr", "synd"), ("dir_", "dir_lappr", "dir_synd")):
            ok, it, post = dec.decode_batch(g[f"s{si}_{lk}"], g[f"s{si}_{sk}"], int(g["maxiter"]), precision="fp64",
                                            schedule=schedule)
            assert np.array_equal(ok.cpu().numpy(), g[f"s{si}_{mode}ok"]), (si, mode)
            assert np.array_equal(it.cpu().numpy(), g[f"s{si}_{mode}it"]), (si, mode)
            assert rel_err(post.cpu().numpy(), g[f"s{si}_{mode}post"]) <= 1e-9, (si, mode)


# ------------------------------------------------------------------ batches against the oracle
def bpsk_frames(orc, vid, cid, frames, sigmas, seed, flip_every=0):
    rng = np.random.default_rng(seed)
    mat = orc.Matrix(vid, cid)
    n = mat.vnum
    word = rng.integers(0, 2, size=(frames, n)).astype(np.uint8)
    sigma = rng.choice(np.asarray(sigmas, dtype=float), size=(frames, 1))
    llr = 2 * ((1 - 2.0 * word) + sigma * rng.normal(size=word.shape)) / sigma ** 2
    synd = np.array([mat.eval_syndrome(w) for w in word])
    if flip_every:
        synd[::flip_every, :3] ^= 1      # inconsistent syndromes -> frames that never converge
    return word, llr, synd


@pytest.mark.parametrize("schedule", [0, 1, 2])
@pytest.mark.parametrize("lanes", [32, 64])
def test_fp64_batch_matches_oracle_with_refill(qr, orc, schedule, lanes):
    """More frames than lanes, widely different convergence times: every frame equals the oracle."""
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(648, 3, 6, seed=2)
    frames = 150
    _, llr, synd = bpsk_frames(orc, vid, cid, frames, [0.6, 0.8, 0.95], seed=21, flip_every=7)
    odec = orc.Decoder(vid, cid)
    ook, oit, opost = odec.decode_frames(llr, synd, 30)
    dec = qr.Decoder(vid, cid)
    ok, it, post = dec.decode_batch(llr, synd, 30, precision="fp64", lanes=lanes, schedule=schedule)
    ok, it, post = ok.cpu().numpy(), it.cpu().numpy(), post.cpu().numpy()
    agree = (ok == ook) & (it == oit)
    assert agree.mean() >= 0.999, (np.flatnonzero(~agree), it[~agree], oit[~agree])
    errs = np.array([rel_err(post[f], opost[f]) for f in range(frames)])
    assert np.max(errs[agree]) <= 1e-9, errs.max()
    assert len(set(oit.tolist())) > 4 and (ook == 0).any() and (ook == 1).any()
    fi, steps = dec.last_stats("fp64", lanes)
    assert fi == int(oit.sum())


def test_fp64_irregular_generic_degrees(qr, orc):
    """Irregular graph with check degrees that hit both the unrolled (<= 8) and the run-time path."""
    from qamreconciliation import codes
    rng = np.random.default_rng(5)
    vid, cid = codes.irregular_ldpc(600, 120, [2, 3, 4, 9], [0.3, 0.4, 0.2, 0.1], seed=6)   # check degrees ~16-17
    v2, c2 = codes.irregular_ldpc(400, 200, [2, 3, 8], [0.5, 0.4, 0.1], seed=7)            # degrees 5-6
    for (vv, cc) in ((vid, cid), (v2, c2)):
        perm = rng.permutation(vv.size)
        vv, cc = vv[perm], cc[perm]                                                          # unsorted edge list
        _, llr, synd = bpsk_frames(orc, vv, cc, 40, [0.5, 0.7], seed=9, flip_every=9)
        ook, oit, opost = orc.Decoder(vv, cc).decode_frames(llr, synd, 25)
        ok, it, post = qr.Decoder(vv, cc).decode_batch(llr, synd, 25, precision="fp64")
        assert np.array_equal(ok.cpu().numpy(), ook) and np.array_equal(it.cpu().numpy(), oit)
        assert rel_err(post.cpu().numpy(), opost) <= 1e-9
        ok32, it32, post32 = qr.Decoder(vv, cc).decode_batch(llr.astype(np.float32), synd, 25, precision="fp32")
        assert (ok32.cpu().numpy() == ook).mean() >= 0.9


def test_fp32_batch_tracks_oracle(qr, orc):
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(6480, 3, 6, seed=1)
    frames = 96
    word, llr, synd = bpsk_frames(orc, vid, cid, frames, [0.80, 0.84], seed=4)
    ook, oit, opost = orc.Decoder(vid, cid).decode_frames(llr, synd, 50)
    dec = qr.Decoder(vid, cid)
    for schedule in (0, 1, 2):
        ok, it, post = dec.decode_batch(torch.tensor(llr, dtype=torch.float32), synd, 50, precision="fp32", schedule=schedule)
        ok, it, post = ok.cpu().numpy(), it.cpu().numpy(), post.cpu().numpy()
        assert post.dtype == np.float32
        assert (ok == ook).mean() >= 0.97
        both = (ok == 1) & (ook == 1)
        assert np.abs(it[both] - oit[both]).max() <= 2
        ndiff = ((post[both] < 0) != (opost[both] < 0)).sum(axis=1)
        assert ndiff.sum() == 0, (schedule, np.flatnonzero(ndiff), ndiff[ndiff > 0], it[both][ndiff > 0], oit[both][ndiff > 0])
        assert np.array_equal((post[both] < 0).astype(np.uint8), word[both])
        # converged frames satisfy their syndrome (checked on the device) and re-decode in 0 iterations
        assert dec.check_lappr_batch(post[both], synd[both]).cpu().numpy().all()
        ok2, it2, _ = dec.decode_batch(post[both], synd[both], 50, precision="fp32")
        assert ok2.cpu().numpy().all() and (it2.cpu().numpy() == 0).all()


def test_edge_cases(qr, orc):
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(96, 3, 6, seed=3)
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid)
    # empty batch
    ok, it, post = dec.decode_batch(np.zeros((0, 96)), np.zeros((0, 48), dtype=np.uint8), 10)
    assert ok.numel() == 0 and it.numel() == 0 and post.shape == (0, 96)
    assert mat.eval_syndrome_batch(np.zeros((0, 96), dtype=np.uint8)).shape == (0, 48)
    # wrong sizes raise (the reference reads out of bounds instead)
    with pytest.raises(ValueError):
        dec.decode(np.zeros(95), np.zeros(48, dtype=np.uint8), 5)
    with pytest.raises(ValueError):
        dec.decode(np.zeros(96), np.zeros(47, dtype=np.uint8), 5)
    with pytest.raises(IndexError):
        mat.eval_syndrome(np.zeros(90, dtype=np.uint8))
    with pytest.raises(ValueError):
        qr.Decoder(np.array([0, 1, 2]), np.array([0, 0]))
    with pytest.raises(ValueError):
        qr.Decoder(np.array([0, 1, 2]), np.array([0, 0, 1]))          # degree-1 check
    with pytest.raises(ValueError):
        qr.NoiseMapper(qr.PAMAlphabet(2, 2), 0.0)
    with pytest.raises(ValueError):
        qr.NoiseMapper(qr.PAMAlphabet(2, 2), 1.0, np.array([0, 1], dtype=np.uint8))
    # max_iterations = 0 and read-only inputs
    _, llr, synd = bpsk_frames(orc, vid, cid, 5, [0.9], seed=2)
    llr.setflags(write=False); synd.setflags(write=False)
    ok, it, post = dec.decode_batch(llr, synd, 0)
    ook, oit, opost = orc.Decoder(vid, cid).decode_frames(llr, synd, 0)
    assert np.array_equal(ok.cpu().numpy(), ook) and np.array_equal(it.cpu().numpy(), oit)
    assert np.array_equal(post.cpu().numpy(), opost)
    # non-finite LLRs (8-PAM hard reverse at high SNR in the reference): no convergence, NaN posteriors
    g = np.load(os.path.join(GOLDEN, "chain_reg_n648_bps3.npz"))
    d3 = qr.Decoder(g["vid"], g["cid"])
    ok, it, post = d3.decode_batch(g["s1_hard_lappr"], g["s1_synd"], 50, precision="fp64")
    assert np.array_equal(ok.cpu().numpy(), g["s1_hard_ok"]) and np.array_equal(it.cpu().numpy(), g["s1_hard_it"])
    assert np.isnan(post.cpu().numpy()).all()
    ok, it, post = d3.decode_batch(g["s1_hard_lappr"], g["s1_synd"], 50, precision="fp32")   # saturating: finite
    assert np.isfinite(post.cpu().numpy()).all()
    # eval_syndrome XORs whole bytes like the reference
    w = np.random.default_rng(0).integers(0, 256, size=(3, 96)).astype(np.uint8)
    want = np.array([orc.Matrix(vid, cid).eval_syndrome(r) for r in w])
    assert np.array_equal(mat.eval_syndrome_batch(w).cpu().numpy(), want)
    # hard decision ties go up, outer thresholds at +-100 a_edge (noisemapper.pyx:41-42, alphabet.pyx:72-73)
    nm = qr.NoiseMapper(qr.PAMAlphabet(2, 2), 1.0)
    assert list(nm.hard_decide_index(np.array([2.0, -2.0, 0.0, 300.0, -300.0, 301.0, -301.0, np.nan]))) == \
        list(orc.NoiseMapper(orc.PAMAlphabet(2, 2), 1.0).hard_decide_index(
            np.array([2.0, -2.0, 0.0, 300.0, -300.0, 301.0, -301.0, np.nan])))


def test_count_errors_and_utils(qr, orc):
    from qamreconciliation import utils
    rng = np.random.default_rng(1)
    lap = rng.normal(size=(7, 500)); lap[0, :5] = 0.0; lap[1, :3] = -0.0
    word = rng.integers(0, 2, size=(7, 500)).astype(np.uint8)
    got = utils.count_errors_batch(lap, word, k=321).cpu().numpy()
    want = [orc.count_errors_from_lappr(lap[f, :321], word[f, :321]) for f in range(7)]
    assert list(got) == want
    assert utils.count_errors_from_lappr(lap[2], word[2]) == orc.count_errors_from_lappr(lap[2], word[2])
    assert list(utils.count_errors_batch(lap.astype(np.float32), word).cpu().numpy()) == \
        [orc.count_errors_from_lappr(lap[f].astype(np.float32).astype(np.float64), word[f]) for f in range(7)]
    with pytest.raises(ValueError):
        utils.count_errors_from_lappr(lap[0], word[0][:-1])
    assert utils.dist_cut(-1) == 0 and utils.dist_cut(2) == 1 and utils.dist_cut(0.3) == 0.3


def test_full_size_properties(qr):
    """BASELINE config 2 size (N = 64 800, E = 194 400): size-independent properties on a batch."""
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(64800, 3, 6, seed=1)
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid)
    pa = qr.PAMAlphabet(2, 2)
    cfg = np.array([0, 1, 0, 1], dtype=np.uint8)
    frames = 80
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    a = torch.tensor(pa.constellation, device="cuda")
    for snr_db, expect_ok in ((5.0, True), (2.0, False)):
        n0 = pa.variance * 10 ** (-snr_db / 10) / 2
        nm = qr.NoiseMapper(pa, n0, cfg)
        x = torch.randint(0, 4, (frames, 32400), device="cuda", generator=gen)
        y = a[x] + np.sqrt(n0) * torch.randn(x.shape, device="cuda", dtype=torch.float64, generator=gen)
        idx, nh, word = nm.front_end_batch(y)
        assert (nh >= 0).all() and (nh <= 1).all()
        synd = mat.eval_syndrome_batch(word)
        # linearity of the syndrome: s(a ^ b) = s(a) ^ s(b)
        other = torch.randint(0, 2, word.shape, device="cuda", dtype=torch.uint8, generator=gen)
        assert torch.equal(mat.eval_syndrome_batch(word ^ other), synd ^ mat.eval_syndrome_batch(other))
        llr = nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float32)
        ok, it, post = dec.decode_batch(llr, synd, 50, precision="fp32")
        assert bool(ok.all()) == expect_ok or not expect_ok
        if expect_ok:
            assert ok.all() and it.max() < 30
            # decoded words are Bob's words and satisfy the syndrome; re-decoding is a no-op
            assert torch.equal((post < 0).to(torch.uint8), word)
            assert dec.check_lappr_batch(post, synd).all()
            ok2, it2, post2 = dec.decode_batch(post, synd, 50, precision="fp32")
            assert ok2.all() and (it2 == 0).all() and torch.equal(post2, post)
        else:
            assert (it[ok == 0] == 50).all() and (ok == 0).float().mean() > 0.9
            # failed frames report the check state truthfully
            assert torch.equal(dec.check_lappr_batch(post, synd), ok)


def test_reconcile_host_matches_stepwise(qr, orc):
    """qr_reconcile_host (host buffers in, host buffers out) == the step-by-step chain."""
    import ctypes as C
    from qamreconciliation import _abi, codes
    vid, cid = codes.regular_ldpc(1296, 3, 6, seed=1)
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid); pa = qr.PAMAlphabet(2, 2)
    cfg = np.array([0, 1, 0, 1], dtype=np.uint8)
    n0 = pa.variance * 10 ** (-7.0 / 10) / 2
    nm = qr.NoiseMapper(pa, n0, cfg)
    rng = np.random.default_rng(3)
    frames, N, K = 70, 1296, 648
    x = rng.integers(0, 4, size=(frames, N // 2)).astype(np.int64)
    y = pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)
    for mode in (0, 1, 2):
        h = dec._handle(_abi.QR_F64, 32)
        ok = np.zeros(frames, np.uint8); it = np.zeros(frames, np.int32); post = np.zeros((frames, N))
        word = np.zeros((frames, N), np.uint8); errs = np.zeros(frames, np.int32)
        _abi.check(_abi.lib().qr_reconcile_host(h, nm._h, mode, 0, 1.0, y.ctypes.data, x.ctypes.data, frames, 50, K,
                                                ok.ctypes.data, it.ctypes.data, post.ctypes.data, _abi.QR_F64,
                                                word.ctypes.data, errs.ctypes.data, None))
        if mode == 2:
            w = pa.demap_symbols_to_bits_batch(x); llr = nm.direct_llr_batch(y)
        else:
            idx, nh, w = nm.front_end_batch(y)
            llr = nm.demap_lappr_array_batch(nh, x) if mode == 0 else nm.bare_llr_batch(x)
        s = mat.eval_syndrome_batch(w)
        ok2, it2, post2 = dec.decode_batch(llr, s, 50, precision="fp64", lanes=32)
        assert np.array_equal(word, w.cpu().numpy())
        assert np.array_equal(ok, ok2.cpu().numpy()) and np.array_equal(it, it2.cpu().numpy())
        assert np.array_equal(post, post2.cpu().numpy())
        want = [orc.count_errors_from_lappr(post[f, :K], word[f, :K]) for f in range(frames)]
        assert list(errs) == want


@pytest.mark.parametrize("n", [96, 100, 1297])
def test_syndrome_of_many_frames_uses_the_staged_kernel(n):
    """>= 256 frames take k_eval_syndrome_smem (one CTA per frame, word packed to bits in shared memory); results must
    be those of the per-frame oracle (matrix.pyx:55-60) for frame lengths that are / are not multiples of 8 (aligned
    and unaligned rows), and for words holding bytes other than 0 / 1 (XOR of whole bytes)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes
    from oracle import port as orc
    rng = np.random.default_rng(n)
    vid, cid = codes.irregular_ldpc(n, n // 2 + 3, [2, 3, 6], [0.3, 0.5, 0.2], seed=n)
    mat, omat = qr.Matrix(vid, cid), orc.Matrix(vid, cid)
    frames = 300
    w = rng.integers(0, 2, size=(frames, n)).astype(np.uint8)
    got = mat.eval_syndrome_batch(w).cpu().numpy()
    want = np.array([omat.eval_syndrome(r) for r in w])
    assert np.array_equal(got, want)
    w[7] = rng.integers(0, 256, size=n).astype(np.uint8)          # one frame of arbitrary bytes among bit frames
    w[299, n - 1] = 2
    got = mat.eval_syndrome_batch(w).cpu().numpy()
    want = np.array([omat.eval_syndrome(r) for r in w])
    assert np.array_equal(got, want)
