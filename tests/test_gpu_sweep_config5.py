"""BASELINE config 5 as a parity case: soft reverse, hard reverse (--hard) and soft direct (--direct)
reconciliation over a 16-point SNR sweep, the GPU's fp32 fast mode against the CPU oracle
(reference drivers: sims/reconciliation.pyx:93-168, :173-249, :253-329).

The north star's bar for the fp32 mode is statistical: BER / FER curves inside the Monte-Carlo confidence
intervals of the reference's.  Two comparisons per SNR point and mode, on a (3,6)-regular n = 6480 code:

  paired    the oracle (fp64, exact bisection demapper -- the reference's arithmetic) and the GPU (fp32 messages,
            fast demapper, fused schedule) decode THE SAME 48 noisy frames: frame-error decisions must agree on
            >= 95 % of the frames of every point and >= 99 % over the sweep, BER within 25 % (+1e-4), average
            iterations over successful frames within 1.5;
  unpaired  an independent GPU run of 1536 frames per point: its FER must lie inside the oracle's Clopper-Pearson
            interval -- the 95 % interval at >= 14 of the 16 points (a 95 % interval misses one time in twenty by
            construction) and the Bonferroni-corrected one (level 1 - 0.05/48: family-wise 95 % over the 48 points
            of the three sweeps) at every point.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

N, BPS, MAXITER = 6480, 2, 50
F_CPU, F_GPU = 48, 1536
SWEEPS = {                                    # the grids of profiles/r1_config5_sweeps (16 points each)
    0: np.linspace(3.0, 6.0, 16),             # soft reverse
    1: np.linspace(4.0, 9.0, 16),             # hard reverse
    2: np.linspace(3.0, 6.0, 16),             # soft direct
}


def clopper_pearson(k, n, level):
    from scipy.stats import beta
    a = (1 - level) / 2
    lo = 0.0 if k == 0 else beta.ppf(a, k, n - k + 1)
    hi = 1.0 if k == n else beta.ppf(1 - a, k + 1, n - k)
    return lo, hi


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes
    from oracle import port as orc
    vid, cid = codes.regular_ldpc(N, 3, 6, seed=11)
    return dict(qr=qr, orc=orc, vid=vid, cid=cid, dec=qr.Decoder(vid, cid), pa=qr.PAMAlphabet(BPS, 2),
                cfg=np.array([0, 1, 0, 1], dtype=np.uint8), K=N - N // 2)


def oracle_point(s, mode, n0, x, y):
    """(bit errors in the first K bits, success, iterations) per frame, the reference's loop body on the oracle."""
    orc = s["orc"]
    pa = orc.PAMAlphabet(BPS, 2.0)
    nm = orc.NoiseMapper(pa, n0, s["cfg"] if mode == 0 else None)
    mat = orc.Matrix(s["vid"], s["cid"])

    def one(f):
        dec = orc.Decoder(s["vid"], s["cid"])
        if mode == 2:
            word = pa.demap_symbols_to_bits(x[f]); llr = orc.direct_llr(y[f], pa, 2 * n0)
        else:
            xh = nm.hard_decide_index(y[f]); word = pa.demap_symbols_to_bits(xh)
            llr = nm.demap_lappr_array(nm.map_noise(y[f], xh), x[f]) if mode == 0 else nm.bare_llr(x[f])
        ok, it, post = dec.decode(llr, mat.eval_syndrome(word), MAXITER)
        return orc.count_errors_from_lappr(post[: s["K"]], word[: s["K"]]), int(ok), int(it)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        return np.array(list(pool.map(one, range(len(x)))))


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["soft-reverse", "hard-reverse", "soft-direct"])
def test_sixteen_point_sweep_inside_oracle_confidence_intervals(setup, mode):
    s = setup
    from qamreconciliation.pipeline import Reconciler
    pa, K = s["pa"], s["K"]
    S = N // BPS
    rows, agree_all, frames_all, miss95 = [], 0, 0, 0
    for p, snr in enumerate(SWEEPS[mode]):
        n0 = pa.variance * 10 ** (-snr / 10) / 2
        rng = np.random.default_rng(1000 * mode + p)
        x = rng.integers(0, 4, size=(F_CPU, S)).astype(np.int64)
        y = pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)
        ref = oracle_point(s, mode, n0, x, y)
        nm = s["qr"].NoiseMapper(pa, n0, s["cfg"]) if mode == 0 else s["qr"].NoiseMapper(pa, n0)
        rec = Reconciler(s["dec"], nm, mode=mode, precision="fp32", demap="fast")
        # ---- paired: the same frames
        out = rec.run_device(torch.tensor(y, device="cuda"), torch.tensor(x, device="cuda"), MAXITER, k_info=K,
                             want_post=False)
        g_err = out["bit_errors"].cpu().numpy(); g_ok = out["success"].cpu().numpy(); g_it = out["iters"].cpu().numpy()
        same = (g_err > 0) == (ref[:, 0] > 0)
        assert same.mean() >= 0.95, (snr, same.mean())
        agree_all += int(same.sum()); frames_all += F_CPU
        ber_ref, ber_gpu = ref[:, 0].sum() / (F_CPU * K), g_err.sum() / (F_CPU * K)
        assert abs(ber_gpu - ber_ref) <= 0.25 * ber_ref + 1e-4, (snr, ber_gpu, ber_ref)
        both = (g_ok == 1) & (ref[:, 1] == 1)
        if both.sum() >= 8:
            assert abs(g_it[both].mean() - ref[both, 2].mean()) <= 1.5, (snr, g_it[both].mean(), ref[both, 2].mean())
        # ---- unpaired: an independent, larger GPU sample against the oracle's binomial interval
        gen = torch.Generator(device="cuda"); gen.manual_seed(77 + 1000 * mode + p)
        xg = torch.randint(0, 4, (F_GPU, S), device="cuda", generator=gen)
        yg = torch.tensor(pa.constellation, device="cuda")[xg] + float(np.sqrt(n0)) * torch.randn(
            (F_GPU, S), device="cuda", dtype=torch.float64, generator=gen)
        big = rec.run_device(yg, xg, MAXITER, k_info=K, want_post=False)
        fer_gpu = float((big["bit_errors"] > 0).float().mean())
        k_ref = int((ref[:, 0] > 0).sum())
        lo95, hi95 = clopper_pearson(k_ref, F_CPU, 0.95)
        lob, hib = clopper_pearson(k_ref, F_CPU, 1 - 0.05 / 48)
        assert lob - 1e-12 <= fer_gpu <= hib + 1e-12, (snr, fer_gpu, k_ref, lob, hib)
        miss95 += not (lo95 <= fer_gpu <= hi95)
        rows.append((snr, k_ref / F_CPU, fer_gpu, ber_ref, ber_gpu))
    assert miss95 <= 2, (miss95, rows)
    assert agree_all / frames_all >= 0.99, agree_all / frames_all
    fers = np.array([r[1] for r in rows])
    assert fers[0] == 1.0 and fers[-1] == 0.0, fers   # the sweep crosses the waterfall
    print("\nEsN0dB  FER(oracle,48)  FER(gpu fp32,1536)  BER(oracle)  BER(gpu, same frames)")
    for r in rows:
        print(f"{r[0]:6.2f}  {r[1]:14.4f}  {r[2]:18.4f}  {r[3]:11.3e}  {r[4]:11.3e}")
