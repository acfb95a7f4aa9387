"""CPU emulation of the product's work-item code (tests/emu) against the oracle and the fixtures.

The emulator compiles qr_decode_core.cuh / qr_mapper_core.cuh / qr_graph_build.h with g++ and
replaces the CUDA grid with loops, so these tests pin the lane state machine (continuous batching,
freeze-at-convergence, iteration counts), the CSR/CSC index arithmetic and the restructured
single-message-array schedule WITHOUT a GPU.  In fp64 the emulation shares libm with the oracle, so
it must be bit-identical to it -- which proves the schedule is an exact restatement.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from oracle import port as orc
from emu_build import load

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
F32, F64 = 32, 64


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def same_bits(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))))


def emu_decode(vid, cid, llr, synd, maxiter, precision=F64, lanes=32, generic=0, post_dtype=F64):
    L = load()
    vid = np.ascontiguousarray(vid, dtype=np.int64); cid = np.ascontiguousarray(cid, dtype=np.int64)
    llr = np.ascontiguousarray(llr); synd = np.ascontiguousarray(synd, dtype=np.uint8)
    frames, N = llr.shape
    ok = np.full(frames, 255, dtype=np.uint8); it = np.full(frames, -1, dtype=np.int32)
    post = np.zeros((frames, N), dtype=np.float64 if post_dtype == F64 else np.float32)
    steps = C.c_int64(0)
    rc = L.emu_decode(ptr(vid), ptr(cid), C.c_int64(vid.size), precision, lanes, generic, ptr(llr),
                      F64 if llr.dtype == np.float64 else F32, ptr(synd), C.c_int64(frames), maxiter, ptr(ok),
                      ptr(it), ptr(post), post_dtype, C.byref(steps))
    assert rc == 0, L.emu_last_error()
    return ok, it, post, steps.value


def emu_decode_fused(vid, cid, llr, synd, maxiter, precision=F64, lanes=32, tile_lanes=32, post_dtype=F64,
                     store_post=1, want_shipped=False):
    L = load()
    vid = np.ascontiguousarray(vid, dtype=np.int64); cid = np.ascontiguousarray(cid, dtype=np.int64)
    llr = np.ascontiguousarray(llr); synd = np.ascontiguousarray(synd, dtype=np.uint8)
    frames, N = llr.shape
    ok = np.full(frames, 255, dtype=np.uint8); it = np.full(frames, -1, dtype=np.int32)
    post = np.zeros((frames, N), dtype=np.float64 if post_dtype == F64 else np.float32)
    steps = C.c_int64(0)
    shipped = C.c_int64(0)
    rc = L.emu_decode_fused(ptr(vid), ptr(cid), C.c_int64(vid.size), precision, lanes, tile_lanes, ptr(llr),
                            F64 if llr.dtype == np.float64 else F32, ptr(synd), C.c_int64(frames), maxiter, ptr(ok),
                            ptr(it), ptr(post), post_dtype, C.byref(steps), store_post, C.byref(shipped))
    if rc == -2:
        return None   # graph not eligible for the fused schedule (variable degree != 3 or a check degree > 8)
    assert rc == 0, L.emu_last_error()
    if want_shipped:
        return ok, it, post, steps.value, shipped.value
    return ok, it, post, steps.value


def test_graph_tables_invariants():
    L = load()
    rng = np.random.default_rng(0)
    from qamreconciliation import codes
    for vid, cid in (codes.hamming_7_4(), codes.regular_ldpc(96, 3, 6, seed=3),
                     codes.irregular_ldpc(300, 150, [2, 3, 8], [0.5, 0.4, 0.1], seed=4)):
        perm = rng.permutation(vid.size)
        vid, cid = np.ascontiguousarray(vid[perm]), np.ascontiguousarray(cid[perm])
        E = vid.size; N = int(vid.max()) + 1; Cn = int(cid.max()) + 1
        dims = np.zeros(6, dtype=np.int64)
        chk_order = np.zeros(Cn, np.int32); slot_edge = np.zeros(E, np.int32); slot_var = np.zeros(E, np.int32)
        var_ptr = np.zeros(N + 1, np.int32); var_slot = np.zeros(E, np.int32)
        assert L.emu_graph_tables(ptr(vid), ptr(cid), C.c_int64(E), ptr(dims), ptr(chk_order), ptr(slot_edge),
                                  ptr(slot_var), ptr(var_ptr), ptr(var_slot)) == 0
        assert list(dims[:3]) == [N, Cn, E]
        assert sorted(chk_order) == list(range(Cn))
        assert sorted(slot_edge) == list(range(E))
        cdeg = np.bincount(cid, minlength=Cn)
        assert np.all(np.diff(cdeg[chk_order]) >= 0)            # internal order sorted by degree
        assert np.array_equal(slot_var, vid[slot_edge])
        # slots of a check are contiguous, in ascending original edge id
        pos = 0
        for c in chk_order:
            d = cdeg[c]
            ed = slot_edge[pos:pos + d]
            assert np.all(cid[ed] == c) and np.all(np.diff(ed) > 0)
            pos += d
        # per-variable lists: ascending original edge id
        for v in range(N):
            ed = slot_edge[var_slot[var_ptr[v]:var_ptr[v + 1]]]
            assert np.all(vid[ed] == v) and np.all(np.diff(ed) > 0)
            assert ed.size == np.count_nonzero(vid == v)


def test_graph_rejects_bad_input():
    L = load()
    dims = np.zeros(6, dtype=np.int64)
    vid = np.array([0, 1, 2], dtype=np.int64); cid = np.array([0, 0, 1], dtype=np.int64)   # check 1 has degree 1
    assert L.emu_graph_tables(ptr(vid), ptr(cid), C.c_int64(3), ptr(dims), None, None, None, None, None) == 2
    assert b"degree" in L.emu_last_error()
    cid = np.array([0, 0, 2], dtype=np.int64)                                               # check 1 missing
    assert L.emu_graph_tables(ptr(vid), ptr(cid), C.c_int64(3), ptr(dims), None, None, None, None, None) == 2
    vid = np.array([0, -1, 2], dtype=np.int64); cid = np.array([0, 0, 0], dtype=np.int64)
    assert L.emu_graph_tables(ptr(vid), ptr(cid), C.c_int64(3), ptr(dims), None, None, None, None, None) == 2


def test_emulated_fp64_decoder_hamming_bit_exact():
    g = np.load(os.path.join(GOLDEN, "decoder_hamming.npz"))
    for name in ("kat1", "kat2", "noconv", "zeros"):
        for mi in (0, 1, 2, 20, 50):
            ok, it, post, _ = emu_decode(g["vid"], g["cid"], g[f"{name}_llr"][None, :], g[f"{name}_synd"][None, :], mi)
            assert ok[0] == int(g[f"{name}_m{mi}_ok"]) and it[0] == int(g[f"{name}_m{mi}_it"]), (name, mi)
            assert same_bits(post[0], g[f"{name}_m{mi}_post"]), (name, mi)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "chain_*.npz"))))
@pytest.mark.parametrize("generic", [0, 1])
def test_emulated_fp64_decoder_chain_bit_exact(path, generic):
    g = np.load(path)
    for si in range(len(g["snrs"])):
        for mode, lk, sk in (("", "lappr", "synd"), ("hard_", "hard_lappr", "synd"), ("dir_", "dir_lappr", "dir_synd")):
            llr = g[f"s{si}_{lk}"]; synd = g[f"s{si}_{sk}"]
            if not np.all(np.isfinite(llr)):
                continue   # +-inf inputs: NaN propagation order differs, covered separately
            ok, it, post, _ = emu_decode(g["vid"], g["cid"], llr, synd, int(g["maxiter"]), generic=generic)
            assert np.array_equal(ok, g[f"s{si}_{mode}ok"]) and np.array_equal(it, g[f"s{si}_{mode}it"])
            assert same_bits(post, g[f"s{si}_{mode}post"])


def test_emulated_fused_schedule_chain_bit_exact():
    """QR_SCHED_FUSED (no stored posteriors, two message buffers, tile-major lanes) restates the same
    arithmetic: fp64 results bit-identical to the compiled reference on every eligible fixture."""
    seen = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "chain_*.npz"))):
        g = np.load(path)
        for si in range(len(g["snrs"])):
            for mode, lk, sk in (("", "lappr", "synd"), ("hard_", "hard_lappr", "synd"), ("dir_", "dir_lappr", "dir_synd")):
                llr = g[f"s{si}_{lk}"]; synd = g[f"s{si}_{sk}"]
                if not np.all(np.isfinite(llr)):
                    continue
                for lanes, tl, sp in ((32, 32, 1), (64, 32, 0), (64, 64, 1)):
                    r = emu_decode_fused(g["vid"], g["cid"], llr, synd, int(g["maxiter"]), lanes=lanes, tile_lanes=tl,
                                         store_post=sp)
                    if r is None:
                        continue
                    ok, it, post, _ = r
                    assert np.array_equal(ok, g[f"s{si}_{mode}ok"]) and np.array_equal(it, g[f"s{si}_{mode}it"])
                    assert same_bits(post, g[f"s{si}_{mode}post"])
                    seen += 1
    assert seen >= 9


def test_emulated_fused_schedule_edge_cases_match_two_phase():
    """max_iterations 0/1, already-consistent input, -0.0, more frames than lanes: fused == two-phase == oracle."""
    from qamreconciliation import codes
    rng = np.random.default_rng(11)
    vid, cid = codes.regular_ldpc(96, 3, 6, seed=5)
    dec = orc.Decoder(vid, cid); mat = orc.Matrix(vid, cid)
    frames = 75
    word = rng.integers(0, 2, size=(frames, 96)).astype(np.uint8)
    sigma = rng.choice([0.5, 0.8, 1.1], size=(frames, 1))
    llr = 2 * ((1 - 2.0 * word) + sigma * rng.normal(size=word.shape)) / sigma ** 2
    synd = np.array([mat.eval_syndrome(w) for w in word])
    synd[::3] ^= rng.integers(0, 2, size=synd[::3].shape).astype(np.uint8)
    llr[5] = np.where(word[5] == 1, -3.0, 3.0); synd[5] = mat.eval_syndrome(word[5])
    llr[7, :4] = -0.0
    for maxiter in (0, 1, 12):
        want = [dec.decode(llr[f], synd[f], maxiter) for f in range(frames)]
        for lanes, tl, sp in ((32, 32, 1), (64, 32, 1), (128, 64, 1), (32, 32, 0)):
            ok, it, post, _, shipped = emu_decode_fused(vid, cid, llr, synd, maxiter, lanes=lanes, tile_lanes=tl,
                                                        store_post=sp, want_shipped=True)
            assert [int(o) for o in ok] == [w[0] for w in want]
            assert [int(i) for i in it] == [w[1] for w in want]
            for f in range(frames):
                assert same_bits(post[f], want[f][2]), (maxiter, lanes, f)
            # posteriors stored by the phase serve the frames that finish at the iteration limit or no earlier
            # than one below the earliest success seen so far; the others are rebuilt -- same bits either way
            iterated = sum(1 for w in want if w[1] > 0)
            if sp == 0:
                assert shipped == 0
            elif maxiter > 0:
                assert 0 < shipped <= iterated
                if maxiter == 12 and lanes == 32:
                    assert shipped >= iterated // 2 and shipped < iterated    # both paths exercised
    # fp32: same decisions as the two-phase fp32 schedule, frame by frame
    ok2, it2, post2, _ = emu_decode(vid, cid, llr.astype(np.float32), synd, 12, precision=F32, post_dtype=F32)
    ok3, it3, post3, _ = emu_decode_fused(vid, cid, llr.astype(np.float32), synd, 12, precision=F32, lanes=64,
                                          tile_lanes=32, post_dtype=F32)
    assert np.array_equal(ok2, ok3) and np.array_equal(it2, it3)
    np.testing.assert_allclose(post3, post2, rtol=1e-5, atol=1e-5)


def test_emulated_continuous_batching_matches_oracle():
    """More frames than lanes, mixed convergence times: every frame must equal the one-at-a-time oracle."""
    from qamreconciliation import codes
    rng = np.random.default_rng(7)
    vid, cid = codes.regular_ldpc(96, 3, 6, seed=5)
    dec = orc.Decoder(vid, cid); mat = orc.Matrix(vid, cid)
    frames = 75
    word = rng.integers(0, 2, size=(frames, 96)).astype(np.uint8)
    sigma = rng.choice([0.5, 0.8, 1.1], size=(frames, 1))
    y = (1 - 2.0 * word) + sigma * rng.normal(size=word.shape)
    llr = 2 * y / sigma ** 2
    # syndrome of a DIFFERENT word for a third of the frames -> some never converge
    synd = np.array([mat.eval_syndrome(w) for w in word])
    synd[::3] ^= rng.integers(0, 2, size=synd[::3].shape).astype(np.uint8)
    llr[5] = np.where(word[5] == 1, -3.0, 3.0)   # already consistent with its own syndrome -> 0 iterations
    synd[5] = mat.eval_syndrome(word[5])
    want = [dec.decode(llr[f], synd[f], 12) for f in range(frames)]
    for lanes in (32, 64):
        ok, it, post, steps = emu_decode(vid, cid, llr, synd, 12, lanes=lanes)
        assert [int(o) for o in ok] == [w[0] for w in want]
        assert [int(i) for i in it] == [w[1] for w in want]
        for f in range(frames):
            assert same_bits(post[f], want[f][2]), f
    assert it[5] == 0 and ok[5] == 1
    assert len(set(int(i) for i in it)) > 3


def test_emulated_fp32_decoder_tracks_oracle():
    from qamreconciliation import codes
    rng = np.random.default_rng(8)
    vid, cid = codes.regular_ldpc(648, 3, 6, seed=2)
    dec = orc.Decoder(vid, cid); mat = orc.Matrix(vid, cid)
    frames = 40
    word = rng.integers(0, 2, size=(frames, 648)).astype(np.uint8)
    sigma = 0.78
    y = (1 - 2.0 * word) + sigma * rng.normal(size=word.shape)
    llr = 2 * y / sigma ** 2
    synd = np.array([mat.eval_syndrome(w) for w in word])
    want = [dec.decode(llr[f], synd[f], 50) for f in range(frames)]
    ok, it, post, _ = emu_decode(vid, cid, llr.astype(np.float32), synd, 50, precision=F32, post_dtype=F32)
    agree = sum(int(ok[f]) == want[f][0] and abs(int(it[f]) - want[f][1]) <= 1 for f in range(frames))
    assert agree >= frames - 2
    for f in range(frames):
        if ok[f] and want[f][0]:
            assert np.array_equal(post[f] < 0, want[f][2] < 0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "mapper_*.npz"))))
def test_emulated_mapper_against_reference(path):
    L = load()
    g = np.load(path)
    bps = int(g["bps"]); M = 1 << bps
    a = np.ascontiguousarray(g["constellation"]); thr = np.ascontiguousarray(g["thresholds"])
    p = np.ascontiguousarray(g["probabilities"]); sc = np.ascontiguousarray(g["sign_config"])
    nv = C.c_double(float(g["noise_var"]))
    y = np.ascontiguousarray(g["y"]); n = y.size
    idx = np.zeros(n, np.int64); nh = np.zeros(n); bits = np.zeros(n * bps, np.uint8)
    L.emu_front(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(y), C.c_int64(n), ptr(idx), ptr(nh), ptr(bits))
    assert np.array_equal(idx, g["idx"]) and np.array_equal(bits, g["bits"])
    np.testing.assert_allclose(nh, g["n_hat"], rtol=0, atol=5e-15)
    for nk, jk, lk in (("n_hat", "x", "lappr"), ("n_grid", "j_grid", "lappr_grid")):
        nn = np.ascontiguousarray(g[nk]); jj = np.ascontiguousarray(g[jk])
        exact = np.zeros(nn.size * bps); fast = np.zeros(nn.size * bps)
        yh_e = np.zeros(nn.size * M); yh_f = np.zeros(nn.size * M)
        L.emu_demap(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 0, ptr(exact), ptr(yh_e))
        L.emu_demap(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 1, ptr(fast), ptr(yh_f))
        # mode bit 2: without the starting table (analytic brackets + float pre-solve): same cells
        fast2 = np.zeros(nn.size * bps); yh_f2 = np.zeros(nn.size * M)
        L.emu_demap(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 5, ptr(fast2), ptr(yh_f2))
        assert np.max(np.abs(yh_f2 - yh_e)) <= 1.1e-9 and np.mean(yh_f2 == yh_e) > 0.98
        # mode bit 3: F table as a starting point only (Halley polish instead of the Hermite solve)
        L.emu_demap(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 9, ptr(fast2), ptr(yh_f2))
        assert np.max(np.abs(yh_f2 - yh_e)) <= 1.1e-9 and np.mean(yh_f2 == yh_e) > 0.98
        # mode bit 4 (16): without the jump table that narrows the grid search -- identical results
        L.emu_demap(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 17, ptr(fast2), ptr(yh_f2))
        assert np.array_equal(yh_f2, yh_f) and np.array_equal(fast2, fast)
        np.testing.assert_allclose(exact, g[lk], rtol=1e-9, atol=1e-9)
        # the fast inverse lands in the same 1e-9 cell as the bisection (a neighbouring one at worst)
        assert np.max(np.abs(yh_f - yh_e)) <= 1.1e-9
        assert np.mean(yh_f == yh_e) > 0.98
        np.testing.assert_allclose(fast, g[lk], rtol=1e-7, atol=1e-7)
    d = np.zeros(n * bps)
    L.emu_direct(bps, ptr(a), C.c_double(2 * float(g["noise_var"])), ptr(y), C.c_int64(n), ptr(d))
    np.testing.assert_allclose(d, g["direct"], rtol=1e-13, atol=1e-13, equal_nan=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "mapper_*.npz"))))
def test_emulated_fp32_grade_demapper_against_reference(path):
    """QR_DEMAP_FAST | QR_DEMAP_F32GRADE (what the fp32 decoder mode is fed; reference noisemapper.pyx:450-559):
    no 2^-30 cell replay, MUFU-grade exp / log / reciprocals.  Gate: 1e-5 relative + 1e-6 absolute against the
    compiled reference's LLRs on the mapper fixtures (thresholds +-1e-12, +-0, +-40, +-1000 included) and on a dense
    random sample against the exact replay."""
    L = load()
    g = np.load(path)
    bps = int(g["bps"]); M = 1 << bps
    a = np.ascontiguousarray(g["constellation"]); thr = np.ascontiguousarray(g["thresholds"])
    p = np.ascontiguousarray(g["probabilities"]); sc = np.ascontiguousarray(g["sign_config"])
    nv = C.c_double(float(g["noise_var"]))
    one = C.c_double(1.0)
    for nk, jk, lk in (("n_hat", "x", "lappr"), ("n_grid", "j_grid", "lappr_grid")):
        nn = np.ascontiguousarray(g[nk]); jj = np.ascontiguousarray(g[jk])
        got = np.zeros(nn.size * bps); fast = np.zeros(nn.size * bps)
        L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 5, one, ptr(got))
        L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 1, one, ptr(fast))
        np.testing.assert_allclose(fast, g[lk], rtol=1e-7, atol=1e-7)          # (demap_symbol's fast path == fixture)
        # a saturated metric (n_hat within 1e-9 of 0 or 1) is defined by where erf rounds: 2 % there, as in exact mode
        sat = np.repeat((nn < 1e-9) | (nn > 1 - 1e-9), bps)
        np.testing.assert_allclose(got[~sat], g[lk][~sat], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got[sat], g[lk][sat], rtol=2e-2, atol=1e-6)
    rng = np.random.default_rng(5)
    nn = np.concatenate([rng.random(6000), 10.0 ** -rng.uniform(1, 12, 600), 1 - 10.0 ** -rng.uniform(1, 12, 600)])
    jj = rng.integers(0, M, size=nn.size).astype(np.int64)
    got = np.zeros(nn.size * bps); exact = np.zeros(nn.size * bps)
    for alpha in (1.0, 0.9):
        L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 5,
                           C.c_double(alpha), ptr(got))
        L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 0,
                           C.c_double(alpha), ptr(exact))
        sat = np.repeat((nn < 1e-9) | (nn > 1 - 1e-9), bps)
        np.testing.assert_allclose(got[~sat], exact[~sat], rtol=1e-5, atol=1e-6)
        assert np.median(np.abs(got - exact) / np.maximum(np.abs(exact), 1e-3)) < 2e-7
    # corrected exponent variant
    L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 7, one, ptr(got))
    L.emu_demap_symbol(bps, ptr(a), ptr(thr), ptr(p), nv, ptr(sc), ptr(nn), ptr(jj), C.c_int64(nn.size), 2, one, ptr(exact))
    sat = np.repeat((nn < 1e-9) | (nn > 1 - 1e-9), bps)
    np.testing.assert_allclose(got[~sat], exact[~sat], rtol=1e-5, atol=1e-6)
