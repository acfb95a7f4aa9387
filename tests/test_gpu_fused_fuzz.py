"""Randomised differential test of the fused flooding-iteration schedule (tile pipeline) against the two-phase
schedule: random (3, k)-regular and variable-regular-3 codes, IRREGULAR variable degrees (1..10, extended neighbour
records), frame counts, lane counts (several tiles, frames >> lanes so that lanes are refilled many times and the
post-processing of one tile overlaps the sweep of the next), tile sizes, iteration limits, noise levels, inconsistent
syndromes.  fp64: success flags, iteration counts and posteriors must be BIT-identical (the two
schedules restate the same arithmetic); fp32: the same flags, iteration counts within 1, same hard decisions on
converged frames.  Catches ordering bugs in the work claims, the bookkeeping and the two shipping paths."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def frames_for(orc, vid, cid, frames, rng):
    mat = orc.Matrix(vid, cid)
    n = mat.vnum
    word = rng.integers(0, 2, size=(frames, n)).astype(np.uint8)
    sigma = rng.choice([0.55, 0.75, 0.9, 1.2], size=(frames, 1))
    llr = 2 * ((1 - 2.0 * word) + sigma * rng.normal(size=word.shape)) / sigma ** 2
    synd = np.array([mat.eval_syndrome(w) for w in word])
    bad = rng.random(frames) < 0.2
    synd[bad, :2] ^= 1                                   # frames that cannot converge
    easy = rng.random(frames) < 0.1
    llr[easy] = np.where(word[easy] == 1, -4.0, 4.0)     # frames that are consistent from the start (0 iterations)
    synd[easy] = np.array([mat.eval_syndrome(w) for w in word[easy]]) if easy.any() else synd[easy]
    return llr, synd


def mixed_check_degrees(codes, n, rng):
    """every variable of degree 3, check degrees drawn from 4..8 (the generic fused kernel, several degree bins)"""
    sockets = 3 * n
    degs = []
    while sum(degs) < sockets:
        degs.append(int(rng.integers(4, 9)))
    over = sum(degs) - sockets
    while over > 0:                                   # trim without leaving the 4..8 range
        i = int(rng.integers(0, len(degs)))
        if degs[i] > 4:
            degs[i] -= 1; over -= 1
    vsock = np.repeat(np.arange(n), 3); rng.shuffle(vsock)
    csock = np.repeat(np.arange(len(degs)), degs)
    vsock = codes._repair_duplicates(vsock, csock, rng)
    return codes._finish(vsock, csock)


def irregular_variable_degrees(codes, n, rng):
    """variable degrees drawn from {1, 2, 3, 4, 5, 10} (inline and extended neighbour records), check degrees <= 8"""
    degs = [1, 2, 3, 4, 5, 10]
    frac = rng.dirichlet(np.ones(len(degs)))
    frac[2] += 1.0; frac /= frac.sum()
    avg = float(np.dot(frac, degs))
    c = int(np.ceil(n * avg / 6.0)) + 2                # average check degree ~6 (max <= 8 for these sizes)
    return codes.irregular_ldpc(n, c, degs, list(frac), seed=int(rng.integers(1, 1 << 30)))


@pytest.mark.parametrize("seed", range(20))
def test_fused_equals_two_phase(seed, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes
    from oracle import port as orc
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([96, 240, 648, 1296]))
    kind = seed % 5
    if kind >= 3:
        vid, cid = irregular_variable_degrees(codes, n, rng)               # any variable degree, mixed check degrees
        if np.bincount(cid).max() > 8:
            pytest.skip("drew a check of degree > 8")
    elif kind == 0:
        vid, cid = codes.regular_ldpc(n, 3, 6, seed=seed)                 # check-regular: the specialised kernel
    elif kind == 1:
        vid, cid = codes.regular_ldpc(n, 3, 3, seed=seed)                 # (3,3): generic fused path, degree 3
    else:
        vid, cid = mixed_check_degrees(codes, n, rng)                      # variable degree 3, check degrees 4..8
    frames = int(rng.integers(1, 220))
    lanes = int(rng.choice([32, 64, 96, 128]))
    maxiter = int(rng.choice([0, 1, 3, 10, 25]))
    llr, synd = frames_for(orc, vid, cid, frames, rng)
    if seed % 4 == 3:
        monkeypatch.setenv("QAMRECON_FUSED_TILE", "64")                    # wider tiles
    if seed % 2:
        monkeypatch.setenv("QAMRECON_FUSED_PP_LAG", str(int(rng.integers(0, 4))))
        monkeypatch.setenv("QAMRECON_FUSED_PP_ITEMS", str(int(rng.integers(1, 9))))
    dec = qr.Decoder(vid, cid)
    assert dec.fused_eligible
    a = dec.decode_batch(llr, synd, maxiter, precision="fp64", lanes=lanes, schedule=0)
    b = dec.decode_batch(llr, synd, maxiter, precision="fp64", lanes=lanes, schedule=2)
    assert torch.equal(a[0], b[0]), (seed, "success")
    assert torch.equal(a[1], b[1]), (seed, "iterations")
    assert torch.equal(a[2].view(torch.int64), b[2].view(torch.int64)), (seed, "posteriors")
    l32 = torch.tensor(llr, dtype=torch.float32)
    a32 = dec.decode_batch(l32, synd, maxiter, precision="fp32", lanes=lanes, schedule=0)
    b32 = dec.decode_batch(l32, synd, maxiter, precision="fp32", lanes=lanes, schedule=2)
    assert torch.equal(a32[0], b32[0]), (seed, "fp32 success")
    assert int((a32[1] - b32[1]).abs().max()) <= 1
    conv = a32[0].bool()
    assert torch.equal(a32[2][conv] < 0, b32[2][conv] < 0)
    # the oracle agrees with both on a few frames
    odec = orc.Decoder(vid, cid)
    for f in range(min(frames, 4)):
        ok, it, post = odec.decode(llr[f], synd[f], maxiter)
        assert (int(a[0][f]), int(a[1][f])) == (ok, it)
        np.testing.assert_allclose(a[2][f].cpu().numpy(), post, rtol=1e-9, atol=1e-9)


def test_default_lane_count_follows_the_batch(monkeypatch):
    """lanes=None: one lane per frame (powers of two from 512 up to 4096, the workspace only grows); the results are
    those of any explicit lane count, bit for bit in fp64 -- with refills (64 lanes) and without (default)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes, _abi
    from oracle import port as orc
    monkeypatch.delenv("QAMRECON_LANES", raising=False)
    rng = np.random.default_rng(77)
    vid, cid = codes.regular_ldpc(240, 3, 6, seed=5)
    dec = qr.Decoder(vid, cid)
    for frames, cap in ((40, 512), (1300, 2048), (700, 2048)):
        llr, synd = frames_for(orc, vid, cid, frames, rng)
        a = dec.decode_batch(llr, synd, 12, precision="fp64", schedule=2)
        assert dec._auto_cap[_abi.QR_F64] == cap
        b = dec.decode_batch(llr, synd, 12, precision="fp64", schedule=2, lanes=64)
        c = dec.decode_batch(llr, synd, 12, precision="fp64", schedule=0)
        for x in (b, c):
            assert torch.equal(a[0], x[0]) and torch.equal(a[1], x[1])
            assert torch.equal(a[2].view(torch.int64), x[2].view(torch.int64))
        fi, _ = dec.last_stats("fp64")
        assert fi == int(torch.where(a[0].bool(), a[1], torch.full_like(a[1], 12)).sum())
