"""The CPU oracle's restatement of the rest of the NoiseMapper surface against fixtures
generated from the compiled reference (tests/golden/make_golden_ext.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import port as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PATHS = sorted(glob.glob(os.path.join(GOLDEN, "mapperext_*.npz")))


def make(g, cls=orc.NoiseMapper):
    pa = orc.PAMAlphabet(int(g["bps"]), 2.0)
    return cls(pa, float(g["noise_var"]), g["sign_config"], float(g["trunkation_threshold"]),
               int(g["n_intervals_per_step"]))


def interior(g):
    """inputs whose interpolation does not sit on the flat tails of the grid (see test below)"""
    n = g["n"]
    return (n > 1e-9) & (n < 1 - 1e-9)


@pytest.mark.parametrize("path", PATHS)
def test_grid_and_cdfs(path):
    g = np.load(path)
    nm = make(g)
    assert nm.y_range.shape == g["y_range"].shape
    assert np.array_equal(nm.y_range, g["y_range"])                 # numpy.linspace reproduced bit for bit
    np.testing.assert_allclose(nm.F_Y_values, g["F_Y_values"], rtol=0, atol=5e-16)   # scipy erf vs libm erf
    np.testing.assert_allclose(nm.F_Y(g["y"]), g["F_Y"], rtol=0, atol=5e-16)
    np.testing.assert_allclose(orc.F_Z(g["y"], 0.7, 1.3), g["F_Z"], rtol=0, atol=3e-16)
    assert np.array_equal(nm.index_to_val(g["idx"]), g["index_to_val"])
    pa = orc.PAMAlphabet(int(g["bps"]), 2.0)
    nd = orc.NoiseMapper(pa, float(g["noise_var"]), g["sign_config"])
    assert nd.y_range.size == int(g["default_grid_points"])
    assert np.array_equal(nd.y_range[[0, 1, -2, -1]], g["default_grid_ends"])


@pytest.mark.parametrize("path", PATHS)
def test_g_inv_and_variants(path):
    g = np.load(path)
    nm = make(g)
    step = float(np.diff(g["y_range"]).max())
    got = nm.demap_noise(g["n"], g["idx"])
    ok = interior(g)
    # where F_Y still moves from grid point to grid point the interpolation is well conditioned ...
    np.testing.assert_allclose(got[ok], g["demap_noise"][ok], rtol=0, atol=1e-9)
    # ... on the flat tails (targets within an ulp of 0 / 1) it picks a grid cell by the last bit of erf
    assert np.all(np.abs(got - g["demap_noise"]) <= 60 * step)
    each = np.array([[nm.g_inv(float(v), i) for i in range(nm.order)] for v in g["n"]])
    np.testing.assert_allclose(each[ok], g["g_inv_each"][ok], rtol=0, atol=1e-9)
    simp = nm.demap_lappr_simplified_array(g["n"], g["idx"]).reshape(-1, nm.bit_per_symbol)
    want = g["simplified"].reshape(-1, nm.bit_per_symbol)
    np.testing.assert_allclose(simp[ok], want[ok], rtol=1e-9, atol=1e-9)
    sof = nm.demap_lappr_sofisticated_array(g["n"], g["idx"]).reshape(-1, nm.bit_per_symbol)
    want = g["sofisticated"].reshape(-1, nm.bit_per_symbol)
    assert np.array_equal(np.isnan(sof[ok]), np.isnan(want[ok]))
    np.testing.assert_allclose(sof[ok], want[ok], rtol=1e-6, atol=1e-6, equal_nan=True)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("cls", ["NoiseMapperFlipSign", "NoiseMapperAntiFlipSign"])
def test_sign_subclasses(path, cls):
    g = np.load(path)
    nm = make(g, getattr(orc, cls))
    ok = interior(g)
    np.testing.assert_allclose(nm.map_noise(g["y"], g["y_idx"]), g[f"{cls}_map_noise"], rtol=0, atol=1e-14)
    np.testing.assert_allclose([nm.g(v, i) for v, i in zip(g["y"], g["y_idx"])], g[f"{cls}_g"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(nm.demap_noise(g["n"], g["idx"])[ok], g[f"{cls}_demap_noise"][ok], rtol=0, atol=1e-9)
    # demap_lappr_array goes through g_inv_search, which the subclasses do not override: ctor sign_config rules
    np.testing.assert_allclose(nm.demap_lappr_array(g["n"][:8].clip(0, 1), g["idx"][:8]), g[f"{cls}_lappr"],
                               rtol=1e-9, atol=1e-9)
    simp = nm.demap_lappr_simplified_array(g["n"], g["idx"]).reshape(-1, nm.bit_per_symbol)
    np.testing.assert_allclose(simp[ok], g[f"{cls}_simplified"].reshape(-1, nm.bit_per_symbol)[ok], rtol=1e-9, atol=1e-9)
