#!/usr/bin/env python3
"""Generate tests/golden/mi_*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref): P_xhat and
montecarlo_information (mutual_information.pyx:29-39, :212-300).

The reference draws its samples with numpy's global RNG (alphabet.pyx:79-83, mutual_information.pyx:236-239);
seeding it and repeating the same calls here gives the samples the reference used, which are stored with its
three estimates so the restatement can be checked on identical inputs.

    python tests/golden/make_golden_mi.py      (this container only: needs oracle/_ref)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import qamreconciliation as ref  # noqa: E402
from qamreconciliation import mutual_information as refmi  # noqa: E402


def main():
    for bps, snr, cname in ((1, 2.0, "base"), (2, 4.0, "base"), (2, 6.0, "alt"), (3, 12.0, "alt")):
        M = 1 << bps
        pa = ref.PAMAlphabet(bps, 2.0)
        cfg = np.zeros(M, dtype=np.uint8)
        if cname == "alt":
            cfg[1::2] = 1
        n0 = pa.variance * 10 ** (-snr / 10) / 2
        nm = ref.NoiseMapper(pa, n0, cfg.copy(), 1e-21, 200)
        p_Xhat = np.array(refmi.P_xhat(nm))
        N, seed = 400, 1000 + bps
        np.random.seed(seed)
        got = refmi.montecarlo_information(pa, nm, p_Xhat.copy(), N)
        # the same draws, repeated: choice(order, N, p) then N scalar randn()
        np.random.seed(seed)
        x_ind = np.array(np.random.choice(M, size=N, p=np.asarray(pa.probabilities)), dtype=np.int64)
        y = np.asarray(pa.constellation)[x_ind].copy()
        for p in range(N):
            y[p] += nm.noise_sigma * np.random.randn()
        path = os.path.join(HERE, f"mi_bps{bps}_{cname}.npz")
        np.savez_compressed(path, bps=bps, noise_var=n0, sign_config=cfg, n_intervals_per_step=200,
                            trunkation_threshold=1e-21, p_Xhat=p_Xhat, x_ind=x_ind, y=y, N=N,
                            estimates=np.array(got))
        print(os.path.basename(path), got)


if __name__ == "__main__":
    main()
