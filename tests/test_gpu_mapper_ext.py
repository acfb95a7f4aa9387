"""GPU parity of the rest of the NoiseMapper surface -- F_Y grid, g_inv /
demap_noise, the simplified and sofisticated LLR formulations, F_Y, F_Z, FlipSign / AntiFlipSign -- through
the Python mirror (i.e. the C ABI) against fixtures from the compiled reference
(tests/golden/make_golden_ext.py) and against the CPU oracle on larger random inputs."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PATHS = sorted(glob.glob(os.path.join(GOLDEN, "mapperext_*.npz")))


@pytest.fixture(scope="module")
def qr():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qamreconciliation
    return qamreconciliation


def make(qr, g, cls="NoiseMapper"):
    pa = qr.PAMAlphabet(int(g["bps"]), 2.0)
    return getattr(qr, cls)(pa, float(g["noise_var"]), g["sign_config"], float(g["trunkation_threshold"]),
                            int(g["n_intervals_per_step"]))


def interior(g):
    n = g["n"]
    return (n > 1e-9) & (n < 1 - 1e-9)


@pytest.mark.parametrize("path", PATHS)
def test_grid_and_cdfs(qr, path):
    g = np.load(path)
    nm = make(qr, g)
    assert np.array_equal(nm.y_range, g["y_range"])                  # numpy.linspace, bit for bit
    np.testing.assert_allclose(nm.F_Y_values, g["F_Y_values"], rtol=0, atol=1e-15)   # CUDA erf vs scipy erf
    np.testing.assert_allclose(nm.F_Y(g["y"]), g["F_Y"], rtol=0, atol=1e-15)
    from qamreconciliation import noisemapper
    np.testing.assert_allclose(noisemapper.F_Z(g["y"], 0.7, 1.3), g["F_Z"], rtol=0, atol=1e-15)
    assert np.array_equal(nm.index_to_val(g["idx"]), g["index_to_val"])
    nd = qr.NoiseMapper(qr.PAMAlphabet(int(g["bps"]), 2.0), float(g["noise_var"]), g["sign_config"])
    assert nd.y_range.size == int(g["default_grid_points"])
    assert np.array_equal(nd.y_range[[0, 1, -2, -1]], g["default_grid_ends"])


@pytest.mark.parametrize("path", PATHS)
def test_g_inv_and_variants(qr, path):
    g = np.load(path)
    nm = make(qr, g)
    step = float(np.diff(g["y_range"]).max())
    ok = interior(g)
    got = nm.demap_noise(g["n"], g["idx"])
    np.testing.assert_allclose(got[ok], g["demap_noise"][ok], rtol=0, atol=1e-9)
    # flat tails of the grid: the cell is picked by the last bit of erf (implementation-defined); bounded only
    assert np.all(np.abs(got - g["demap_noise"]) <= 60 * step)
    assert abs(nm.g_inv(float(g["n"][0]), int(g["idx"][0])) - g["demap_noise"][0]) <= 1e-9
    b = nm.bit_per_symbol
    simp = nm.demap_lappr_simplified_array(g["n"], g["idx"]).reshape(-1, b)
    np.testing.assert_allclose(simp[ok], g["simplified"].reshape(-1, b)[ok], rtol=1e-9, atol=1e-9)
    sof = nm.demap_lappr_sofisticated_array(g["n"], g["idx"]).reshape(-1, b)
    want = g["sofisticated"].reshape(-1, b)
    assert np.array_equal(np.isnan(sof[ok]), np.isnan(want[ok]))
    np.testing.assert_allclose(sof[ok], want[ok], rtol=1e-6, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(nm.demap_lappr_simplified(float(g["n"][1]), int(g["idx"][1])), simp[1], rtol=0, atol=0)
    with pytest.raises(ValueError):
        nm.demap_noise(g["n"], g["idx"][:-1])
    with pytest.raises(ValueError):
        nm.demap_lappr_simplified_array(g["n"], g["idx"][:-1])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("cls", ["NoiseMapperFlipSign", "NoiseMapperAntiFlipSign"])
def test_sign_subclasses(qr, path, cls):
    g = np.load(path)
    nm = make(qr, g, cls)
    ok = interior(g)
    b = nm.bit_per_symbol
    np.testing.assert_allclose(nm.map_noise(g["y"], g["y_idx"]), g[f"{cls}_map_noise"], rtol=0, atol=1e-14)
    np.testing.assert_allclose([nm.g(v, i) for v, i in zip(g["y"][:6], g["y_idx"][:6])], g[f"{cls}_g"][:6], rtol=0, atol=1e-14)
    np.testing.assert_allclose(nm.demap_noise(g["n"], g["idx"])[ok], g[f"{cls}_demap_noise"][ok], rtol=0, atol=1e-9)
    # demap_lappr_array on the subclasses uses the constructor's sign_config (g_inv_search is not overridden)
    np.testing.assert_allclose(nm.demap_lappr_array(g["n"][:8].clip(0, 1), g["idx"][:8]), g[f"{cls}_lappr"],
                               rtol=1e-9, atol=1e-9)
    simp = nm.demap_lappr_simplified_array(g["n"], g["idx"]).reshape(-1, b)
    np.testing.assert_allclose(simp[ok], g[f"{cls}_simplified"].reshape(-1, b)[ok], rtol=1e-9, atol=1e-9)


def test_batched_variants_against_oracle(qr):
    """a frame-sized batch (default 1000-intervals-per-step grid): CUDA vs the CPU oracle"""
    from oracle import port as orc
    rng = np.random.default_rng(3)
    bps = 2
    pa = qr.PAMAlphabet(bps, 2.0); opa = orc.PAMAlphabet(bps, 2.0)
    n0 = pa.variance * 10 ** (-4.0 / 10) / 2
    cfg = np.array([0, 1, 0, 1], dtype=np.uint8)
    nm = qr.NoiseMapper(pa, n0, cfg); onm = orc.NoiseMapper(opa, n0, cfg)
    n = rng.uniform(1e-6, 1 - 1e-6, size=(4, 5000)); j = rng.integers(0, 4, size=n.shape)
    yh = nm.demap_noise_batch(n, j).cpu().numpy()
    np.testing.assert_allclose(yh.ravel(), onm.demap_noise(n.ravel(), j.ravel()), rtol=0, atol=1e-9)
    s1 = nm.demap_lappr_simplified_array_batch(n, j).cpu().numpy()
    assert s1.shape == (4, 10000)
    np.testing.assert_allclose(s1.ravel(), onm.demap_lappr_simplified_array(n.ravel(), j.ravel()), rtol=1e-9, atol=1e-9)
    s2 = nm.demap_lappr_sofisticated_array_batch(n, j).cpu().numpy()
    np.testing.assert_allclose(s2.ravel(), onm.demap_lappr_sofisticated_array(n.ravel(), j.ravel()), rtol=1e-6,
                               atol=1e-6, equal_nan=True)
