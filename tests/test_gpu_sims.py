"""GPU tests of the batched reconciliation pass and the Monte-Carlo drivers:
paired comparison with the CPU oracle on identical channel outputs, and BER / FER / iteration
statistics within Monte-Carlo confidence intervals (the fp32 bar of the north star)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import codes
    from oracle import port as orc
    vid, cid = codes.regular_ldpc(648, 3, 6, seed=2)
    return dict(qr=qr, orc=orc, vid=vid, cid=cid, dec=qr.Decoder(vid, cid), mat=qr.Matrix(vid, cid),
                pa=qr.PAMAlphabet(2, 2), odec=orc.Decoder(vid, cid), omat=orc.Matrix(vid, cid),
                opa=orc.PAMAlphabet(2, 2.0), cfg=np.array([0, 1, 0, 1], dtype=np.uint8))


def oracle_frames(s, snr_db, x, y, mode):
    """The reference's loop body (sims/reconciliation.pyx:129-156 / :211-233 / :297-314) on given x, y."""
    orc, pa = s["orc"], s["opa"]
    n0 = pa.variance * 10 ** (-snr_db / 10) / 2
    nm = orc.NoiseMapper(pa, n0, s["cfg"] if mode == 0 else None)
    K = 324
    rows = []
    for f in range(x.shape[0]):
        if mode == 2:
            word = pa.demap_symbols_to_bits(x[f]); llr = orc.direct_llr(y[f], pa, 2 * n0)
        else:
            xh = nm.hard_decide_index(y[f]); word = pa.demap_symbols_to_bits(xh)
            llr = nm.demap_lappr_array(nm.map_noise(y[f], xh), x[f]) if mode == 0 else nm.bare_llr(x[f])
        synd = s["omat"].eval_syndrome(word)
        ok, it, post = s["odec"].decode(llr, synd, 50)
        rows.append((orc.count_errors_from_lappr(post[:K], word[:K]), ok, it))
    return np.array(rows)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_paired_frames_against_oracle(setup, mode):
    s = setup
    from qamreconciliation.pipeline import Reconciler
    snr = {0: 4.4, 1: 6.0, 2: 4.4}[mode]
    rng = np.random.default_rng(10 + mode)
    frames = 160
    pa = s["pa"]
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    x = rng.integers(0, 4, size=(frames, 324)).astype(np.int64)
    y = pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)
    want = oracle_frames(s, snr, x, y, mode)
    nm = s["qr"].NoiseMapper(pa, n0, s["cfg"] if mode == 0 else None)
    # fp64 + exact demapper: the parity mode
    out = Reconciler(s["dec"], nm, mode=mode, precision="fp64", demap="exact").run_device(
        torch.tensor(y, device="cuda"), torch.tensor(x, device="cuda"), 50, k_info=324)
    got = np.stack([out["bit_errors"].cpu().numpy(), out["success"].cpu().numpy(), out["iters"].cpu().numpy()], axis=1)
    agree = np.all(got == want, axis=1)
    assert agree.mean() >= 0.99, np.flatnonzero(~agree)
    # fp32 + fast demapper: same decisions on (nearly) every frame
    out = Reconciler(s["dec"], nm, mode=mode, precision="fp32", demap="fast").run_device(
        torch.tensor(y, device="cuda"), torch.tensor(x, device="cuda"), 50, k_info=324)
    got32 = np.stack([out["bit_errors"].cpu().numpy(), out["success"].cpu().numpy(), out["iters"].cpu().numpy()], axis=1)
    assert (got32[:, 1] == want[:, 1]).mean() >= 0.95
    assert ((got32[:, 0] > 0) == (want[:, 0] > 0)).mean() >= 0.95
    both = (got32[:, 1] == 1) & (want[:, 1] == 1)
    assert np.abs(got32[both, 2] - want[both, 2]).max() <= 3
    assert 0 < want[:, 1].sum() < frames or mode != 0       # the soft-RR point sits in the waterfall


@pytest.mark.parametrize("precision,schedule", [("fp32", "0"), ("fp64", "0"), ("fp32", "3")])
def test_monte_carlo_statistics_within_confidence_intervals(setup, precision, schedule, monkeypatch):
    """BER / FER / average iterations of simulate_softening_snr_dB against the oracle's own Monte Carlo
    (schedule 3 = the fused flooding iteration where the graph allows it)."""
    s = setup
    from sims.reconciliation import simulate_softening_snr_dB
    snr = 4.3
    rng = np.random.default_rng(99)
    n_ref = 500
    pa = s["pa"]
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    x = rng.integers(0, 4, size=(n_ref, 324)).astype(np.int64)
    y = pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)
    ref = oracle_frames(s, snr, x, y, 0)
    fer_ref = (ref[:, 0] > 0).mean(); ber_ref = ref[:, 0].sum() / (n_ref * 324)
    it_ref = ref[ref[:, 1] == 1, 2].mean()
    monkeypatch.setenv("QAMRECON_PRECISION", precision)
    monkeypatch.setenv("QAMRECON_DEMAP", "fast" if precision == "fp32" else "exact")
    monkeypatch.setenv("QAMRECON_SCHEDULE", schedule)
    np.random.seed(5)
    n_gpu = 4000
    snr_out, ber, fer, avg_it = simulate_softening_snr_dB(snr, s["dec"], s["mat"], pa, s["cfg"], 50, n_gpu, 10 ** 9)
    assert snr_out == snr
    sig = np.sqrt(fer_ref * (1 - fer_ref) * (1 / n_ref + 1 / n_gpu))
    assert abs(fer - fer_ref) <= 4 * sig + 1e-3, (fer, fer_ref, sig)
    assert abs(ber - ber_ref) <= 0.35 * max(ber_ref, 1e-4), (ber, ber_ref)
    assert abs(avg_it - it_ref) <= 1.0, (avg_it, it_ref)
    assert 0.1 < fer_ref < 0.7


def test_driver_semantics_and_modes(setup, monkeypatch):
    s = setup
    from sims.reconciliation import (simulate_direct_snr_dB, simulate_hard_reverse_snr_dB,
                                     simulate_softening_snr_dB, y_to_lappr_grey_array)
    np.random.seed(1)
    # high SNR: no errors, few iterations; the loop runs all simulation_loops frames
    r = simulate_softening_snr_dB(8.0, s["dec"], s["mat"], s["pa"], s["cfg"], 50, 300, 100)
    assert r[0] == 8.0 and r[1] == 0 and r[2] == 0 and 0 < r[3] < 6
    # low SNR: every frame fails -> stops right after simulation_loops/20 once ferr_count_min is met
    monkeypatch.setenv("QAMRECON_SIM_BATCH", "64")
    r = simulate_softening_snr_dB(1.0, s["dec"], s["mat"], s["pa"], s["cfg"], 10, 400, 5)
    assert r[2] == 1.0 and r[3] == 0 and 0.02 < r[1] < 0.3
    r_h = simulate_hard_reverse_snr_dB(9.0, s["dec"], s["mat"], s["pa"], 50, 200, 100)
    r_d = simulate_direct_snr_dB(8.0, s["dec"], s["mat"], s["pa"], 50, 200, 100)
    assert r_h[2] < 0.05 and r_d[2] == 0 and r_d[3] > 0
    y = np.array([-2.5, 0.3, 1.9])
    got = y_to_lappr_grey_array(y, s["pa"], 1.7)
    np.testing.assert_allclose(got, s["orc"].direct_llr(y, s["opa"], 1.7), rtol=1e-12, atol=1e-12)
    # N not a multiple of bits per symbol is rejected (the reference raises IndexError much later)
    from qamreconciliation import codes
    hv, hc = codes.hamming_7_4()
    with pytest.raises(ValueError):
        simulate_softening_snr_dB(5.0, s["qr"].Decoder(hv, hc), s["qr"].Matrix(hv, hc), s["pa"], s["cfg"], 5, 10, 1)
    # bps = 1 works on the Hamming code (BASELINE config 1 through the PAM chain)
    pa1 = s["qr"].PAMAlphabet(1, 2)
    r = simulate_softening_snr_dB(9.0, s["qr"].Decoder(hv, hc), s["qr"].Matrix(hv, hc), pa1,
                                  np.array([0, 1], dtype=np.uint8), 20, 200, 1000)
    assert r[2] < 0.2


@pytest.mark.parametrize("precision,demap", [("fp64", "exact"), ("fp32", "fast")])
def test_single_call_chain_equals_stagewise(setup, precision, demap):
    """qr_reconcile_device (whole chain in one library call) == the stages run one library call at a time
    (more frames than lanes, frames finishing at different iterations)."""
    s = setup
    from qamreconciliation.pipeline import Reconciler
    rng = np.random.default_rng(77)
    pa = s["pa"]
    frames = 300
    snr = 4.4
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    x = torch.tensor(rng.integers(0, 4, size=(frames, 324)), device="cuda")
    y = torch.tensor(pa.constellation, device="cuda")[x] + np.sqrt(n0) * torch.tensor(rng.normal(size=(frames, 324)), device="cuda")
    nm = s["qr"].NoiseMapper(pa, n0, s["cfg"])
    for lanes in (32, 96):
        rec = Reconciler(s["dec"], nm, precision=precision, demap=demap, lanes=lanes)
        a = rec.run_device(y, x, 50, k_info=324)
        b = rec.run_device(y, x, 50, k_info=324, stagewise=True)
        for k in ("success", "iters", "word", "synd", "bit_errors"):
            assert torch.equal(a[k], b[k]), k
        assert torch.equal(a["post"].view(torch.uint8), b["post"].view(torch.uint8))
        assert 0 < int(a["success"].sum()) < frames
