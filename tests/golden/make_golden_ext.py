#!/usr/bin/env python3
"""Generate tests/golden/mapperext_*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref):
the rest of the NoiseMapper surface -- the dense F_Y grid, g_inv /
demap_noise (grid interpolation), demap_lappr_simplified / _sofisticated, F_Y, F_Z, index_to_val and the
FlipSign / AntiFlipSign subclasses (noisemapper.pyx:47-98, :264-307, :362-419, :563-816).

    python tests/golden/make_golden_ext.py      (this container only: needs oracle/_ref)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import qamreconciliation as ref  # noqa: E402  (the compiled reference)
from qamreconciliation import noisemapper as refnm  # noqa: E402


def w(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype).copy()


def arr(mv):
    return np.array(np.asarray(mv), copy=True)


def main():
    rng = np.random.default_rng(20261018)
    for bps in (1, 2, 3):
        M = 1 << bps
        pa = ref.PAMAlphabet(bps, 2.0)
        for cname in ("base", "alt"):
            cfg = np.zeros(M, dtype=np.uint8)
            if cname == "alt":
                cfg[1::2] = 1
            snr = 4.0 if bps < 3 else 12.0
            n0 = pa.variance * 10 ** (-snr / 10) / 2
            # coarse grid (n_intervals_per_step=50) keeps the fixture small; the default 1000 is covered by shape only
            nm = ref.NoiseMapper(pa, n0, w(cfg, np.uint8), 1e-21, 50)
            out = dict(bps=bps, noise_var=n0, sign_config=cfg, n_intervals_per_step=50, trunkation_threshold=1e-21,
                       constellation=arr(pa.constellation), thresholds=arr(pa.thresholds),
                       probabilities=arr(pa.probabilities))
            out["y_range"] = nm.y_range
            out["F_Y_values"] = nm.F_Y_values
            nd = ref.NoiseMapper(pa, n0, w(cfg, np.uint8))
            out["default_grid_points"] = np.int64(nd.y_range.size)
            out["default_grid_ends"] = np.array([nd.y_range[0], nd.y_range[1], nd.y_range[-2], nd.y_range[-1]])
            # inputs: interior values, values next to 0 and 1, exact 0 / 1, out-of-range
            n = np.concatenate([rng.uniform(0, 1, 40), [0.0, 1.0, 1e-12, 1 - 1e-12, 1e-6, 1 - 1e-6, 0.5, -0.25, 1.25]])
            idx = rng.integers(0, M, n.size)
            out["n"] = n
            out["idx"] = idx
            out["demap_noise"] = arr(nm.demap_noise(w(n, np.float64), w(idx, np.int64)))
            out["g_inv_each"] = np.array([[nm.g_inv(float(v), i) for i in range(M)] for v in n])
            out["simplified"] = arr(nm.demap_lappr_simplified_array(w(n, np.float64), w(idx, np.int64)))
            out["sofisticated"] = arr(nm.demap_lappr_sofisticated_array(w(n, np.float64), w(idx, np.int64)))
            y = np.concatenate([rng.normal(0, 3.0, 30), [0.0, -0.0, 40.0, -40.0]])
            out["y"] = y
            out["F_Y"] = arr(nm.F_Y(w(y, np.float64)))
            out["F_Z"] = arr(refnm.F_Z(w(y, np.float64), 0.7, 1.3))
            out["index_to_val"] = arr(nm.index_to_val(w(idx, np.int64)))
            yi = arr(nm.hard_decide_index(w(y, np.float64)))
            out["y_idx"] = yi
            for cls in ("NoiseMapperFlipSign", "NoiseMapperAntiFlipSign"):
                sub = getattr(ref, cls)(pa, n0, w(cfg, np.uint8), 1e-21, 50)
                out[f"{cls}_map_noise"] = arr(sub.map_noise(w(y, np.float64), w(yi, np.int64)))
                out[f"{cls}_demap_noise"] = arr(sub.demap_noise(w(n, np.float64), w(idx, np.int64)))
                out[f"{cls}_g"] = np.array([sub.g(float(v), int(i)) for v, i in zip(y, yi)])
                # demap_lappr_array on the subclasses goes through g_inv_search, which they do NOT override
                out[f"{cls}_lappr"] = arr(sub.demap_lappr_array(w(n[:8].clip(0, 1), np.float64), w(idx[:8], np.int64)))
                out[f"{cls}_simplified"] = arr(sub.demap_lappr_simplified_array(w(n, np.float64), w(idx, np.int64)))
            path = os.path.join(HERE, f"mapperext_bps{bps}_{cname}.npz")
            np.savez_compressed(path, **out)
            print(os.path.basename(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
