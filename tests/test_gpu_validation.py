"""Argument validation the reference gets from bounds-checked Cython (IndexError on symbol indices outside the
alphabet; Matrix on any edge list) and the shape / dtype checks of the batched entry points."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    return qr


def test_out_of_range_symbol_indices_raise_index_error(qr):
    pa = qr.PAMAlphabet(2, 2)
    nm = qr.NoiseMapper(pa, 0.5, np.array([0, 1, 0, 1], dtype=np.uint8))
    y = np.array([-2.5, 0.3, 1.9, 2.2])
    good = np.array([0, 2, 2, 3], dtype=np.int64)
    for bad in (np.array([0, 4, 2, 3]), np.array([0, -1, 2, 3]), np.array([0, 1 << 40, 2, 3])):
        bad = bad.astype(np.int64)
        with pytest.raises(IndexError):
            nm.map_noise(y, bad)
        with pytest.raises(IndexError):
            nm.demap_lappr_array(np.array([0.2, 0.4, 0.6, 0.8]), bad)
        with pytest.raises(IndexError):
            nm.bare_llr(bad)
        with pytest.raises(IndexError):
            nm.demap_noise_search(np.array([0.2, 0.4, 0.6, 0.8]), bad)
        with pytest.raises(IndexError):
            nm.demap_noise(np.array([0.2, 0.4, 0.6, 0.8]), bad)
        with pytest.raises(IndexError):
            nm.demap_lappr_simplified_array(np.array([0.2, 0.4, 0.6, 0.8]), bad)
        with pytest.raises(IndexError):
            pa.demap_symbols_to_bits(bad)
    # the counter is reset by the raise: good input afterwards is fine, and the context is not poisoned
    assert nm.map_noise(y, good).shape == (4,)
    assert pa.demap_symbols_to_bits(good).shape == (8,)
    # batched entry points: the offence is reported at the next check_indices()
    nm.bare_llr_batch(torch.tensor([[0, 7]], device="cuda"))
    with pytest.raises(IndexError):
        nm.check_indices()
    nm.check_indices()


def test_reconciler_validates_shapes(qr):
    from qamreconciliation import codes
    from qamreconciliation.pipeline import Reconciler
    vid, cid = codes.regular_ldpc(96, 3, 6, seed=2)
    pa = qr.PAMAlphabet(2, 2)
    nm = qr.NoiseMapper(pa, 0.5, np.array([0, 1, 0, 1], dtype=np.uint8))
    rec = Reconciler(qr.Decoder(vid, cid), nm)
    y = torch.zeros((3, 48), dtype=torch.float64, device="cuda"); x = torch.zeros((3, 48), dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        rec.run_device(y[:, :40], x[:, :40], 5)
    with pytest.raises(ValueError):
        rec.run_device(y, x[:2], 5)
    out = dict(success=torch.empty(3, dtype=torch.uint8), iters=torch.empty(3, dtype=torch.int32),
               bit_errors=torch.empty(3, dtype=torch.int32))
    with pytest.raises(ValueError):
        rec.run_host(y, x.cpu(), 5, 48, out)                       # y on the device
    with pytest.raises(ValueError):
        rec.run_host(y.cpu().float(), x.cpu(), 5, 48, out)         # wrong dtype
    with pytest.raises(ValueError):
        rec.run_host(y.cpu(), x.cpu(), 5, 48, dict(out, iters=torch.empty(2, dtype=torch.int32)))
    x_bad = x.cpu().clone(); x_bad[1, 3] = 9
    with pytest.raises(IndexError):
        rec.run_host(y.cpu(), x_bad, 5, 48, out)
    rec.run_host(y.cpu(), x.cpu(), 5, 48, out)


def test_matrix_takes_any_edge_list_decoder_does_not(qr):
    """matrix.pyx:21-38 accepts checks of degree 0 / 1 and degrees above 64; eval_syndrome works on them."""
    # check 0: degree 1, check 1: unused (degree 0), check 2: degree 70
    vid = np.concatenate([[0], np.arange(70)]).astype(np.int64)
    cid = np.concatenate([[0], np.full(70, 2)]).astype(np.int64)
    mat = qr.Matrix(vid, cid)
    assert (mat.vnum, mat.cnum, mat.ednum) == (70, 3, 71)
    rng = np.random.default_rng(0)
    words = rng.integers(0, 2, size=(5, 70)).astype(np.uint8)
    got = mat.eval_syndrome_batch(words).cpu().numpy()
    want = np.stack([words[:, 0], np.zeros(5, np.uint8), words.sum(axis=1) % 2], axis=1).astype(np.uint8)
    assert np.array_equal(got, want)
    assert np.array_equal(mat.eval_syndrome(words[0]), want[0])
    with pytest.raises(ValueError, match="degree"):
        qr.Decoder(vid, cid)
