"""qr_reconcile_host_compact (float32 samples, byte symbols in; packed hard decisions out) against qr_reconcile_host on
the same data widened on the host: identical flags, iteration counts, bit errors, and decisions == (posterior < 0)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_compact_wire_format_equals_reference_types(mode):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes
    from qamreconciliation.pipeline import Reconciler
    n, frames = 1296, 150
    vid, cid = codes.regular_ldpc(n, 3, 6, seed=3)
    pa = qr.PAMAlphabet(2, 2)
    n0 = pa.variance * 10 ** (-4.6 / 10) / 2
    nm = qr.NoiseMapper(pa, n0, np.array([0, 1, 0, 1], dtype=np.uint8)) if mode == 0 else qr.NoiseMapper(pa, n0)
    rng = np.random.default_rng(8)
    x = rng.integers(0, 4, size=(frames, n // 2))
    y32 = (pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)).astype(np.float32)
    for precision in ("fp32", "fp64"):
        rec = Reconciler(qr.Decoder(vid, cid), nm, mode=mode, precision=precision, demap="fast", lanes=64)
        dt = torch.float32 if precision == "fp32" else torch.float64
        ref = dict(success=torch.empty(frames, dtype=torch.uint8), iters=torch.empty(frames, dtype=torch.int32),
                   bit_errors=torch.empty(frames, dtype=torch.int32), post=torch.empty((frames, n), dtype=dt))
        rec.run_host(torch.tensor(y32.astype(np.float64)), torch.tensor(x.astype(np.int64)), 50, n // 2, ref)
        got = dict(success=torch.empty(frames, dtype=torch.uint8), iters=torch.empty(frames, dtype=torch.int32),
                   bit_errors=torch.empty(frames, dtype=torch.int32),
                   decisions=torch.empty((frames, (n + 7) // 8), dtype=torch.uint8))
        rec.run_host_compact(torch.tensor(y32), torch.tensor(x.astype(np.uint8)), 50, n // 2, got)
        for k in ("success", "iters", "bit_errors"):
            assert torch.equal(ref[k], got[k]), (precision, k)
        bits = np.unpackbits(got["decisions"].numpy(), axis=1, bitorder="little")[:, :n]
        assert np.array_equal(bits, (ref["post"].numpy() < 0).astype(np.uint8))
        assert 0 < int(ref["success"].sum())
    with pytest.raises(ValueError):
        rec.run_host_compact(torch.tensor(y32.astype(np.float64)), torch.tensor(x.astype(np.uint8)), 50, n // 2, got)
