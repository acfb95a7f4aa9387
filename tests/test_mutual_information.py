"""Mutual-information Monte Carlo (mutual_information.pyx:29-39, :212-300):
the CPU oracle and the CUDA kernel against estimates the compiled reference produced on the SAME samples
(tests/golden/make_golden_mi.py re-creates the reference's numpy draws)."""
import glob
import os

import numpy as np
import pytest

from oracle import port as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PATHS = sorted(glob.glob(os.path.join(GOLDEN, "mi_*.npz")))


def make(mod, g):
    pa = mod.PAMAlphabet(int(g["bps"]), 2.0)
    return pa, mod.NoiseMapper(pa, float(g["noise_var"]), g["sign_config"], float(g["trunkation_threshold"]),
                               int(g["n_intervals_per_step"]))


@pytest.mark.parametrize("path", PATHS)
def test_oracle_matches_reference_estimates(path):
    g = np.load(path)
    pa, nm = make(orc, g)
    np.testing.assert_allclose(orc.P_xhat(nm), g["p_Xhat"], rtol=1e-14, atol=0)
    got = orc.information_from_samples(nm, g["p_Xhat"], g["x_ind"], g["y"])
    np.testing.assert_allclose(got, g["estimates"], rtol=1e-12, atol=1e-12)
    only = orc.information_from_samples(nm, g["p_Xhat"], g["x_ind"], g["y"], which=(0, 1, 0))
    assert only[0] == 0 and only[2] == 0 and abs(only[1] - g["estimates"][1]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("path", PATHS)
def test_gpu_matches_reference_estimates(path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import mutual_information as mi
    g = np.load(path)
    pa, nm = make(qr, g)
    np.testing.assert_allclose(mi.P_xhat(nm), g["p_Xhat"], rtol=1e-13, atol=0)
    for mode in ("exact", "fast"):
        got = mi.information_from_samples(nm, g["p_Xhat"], g["x_ind"], g["y"], mode=mode)
        # 400 samples, sums of logs of erf/exp-based terms: CUDA vs libm differ in the last ulps
        np.testing.assert_allclose(got, g["estimates"], rtol=1e-10, atol=1e-10)
    only = mi.information_from_samples(nm, g["p_Xhat"], g["x_ind"], g["y"], which=(1, 0, 0))
    assert only[1] == 0 and only[2] == 0 and abs(only[0] - g["estimates"][0]) < 1e-10


@pytest.mark.gpu
def test_gpu_montecarlo_converges_to_the_oracle_on_large_samples():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation as qr
    from qamreconciliation import mutual_information as mi
    bps, snr = 2, 4.0
    pa = qr.PAMAlphabet(bps, 2.0); opa = orc.PAMAlphabet(bps, 2.0)
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    cfg = np.array([0, 1, 0, 1], dtype=np.uint8)
    nm = qr.NoiseMapper(pa, n0, cfg); onm = orc.NoiseMapper(opa, n0, cfg)
    p_Xhat = mi.P_xhat(nm)
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    big = mi.montecarlo_information(pa, nm, p_Xhat, 2_000_000, generator=gen, mode="fast")
    # the same estimator on 20 000 oracle samples: agreement within Monte-Carlo error (std of a term ~ 1)
    rng = np.random.default_rng(6)
    x = rng.integers(0, 4, 20_000); y = opa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.size)
    small = orc.information_from_samples(onm, p_Xhat, x, y)
    assert np.all(np.abs(np.array(big) - np.array(small)) < 0.05)
    # information-theoretic sanity: I(X;Xhat) <= I(X,N;Xhat) <= I(X;Y) <= bits per symbol (estimates are of
    # -I(X;Xhat), -I(X;Y)... in the reference's sign convention: the first two come out negative)
    assert -big[0] <= big[2] + 0.01 and big[2] <= -big[1] + 0.01 and -big[1] <= bps
