#!/usr/bin/env python3
"""Tiny run of every kernel for compute-sanitizer (memcheck / racecheck): small codes, more frames than lanes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import codes
from qamreconciliation.pipeline import Reconciler

rng = np.random.default_rng(0)
for (vid, cid), bps in ((codes.regular_ldpc(96, 3, 6, seed=3), 2), (codes.irregular_ldpc(300, 150, [2, 3, 8], [0.5, 0.4, 0.1], seed=4), 2),
                        (codes.hamming_7_4(), 1)):
    n = int(vid.max()) + 1
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid); pa = qr.PAMAlphabet(bps, 2)
    cfg = np.zeros(pa.order, dtype=np.uint8); cfg[1::2] = 1
    n0 = pa.variance * 10 ** (-5.0 / 10) / 2
    nm = qr.NoiseMapper(pa, n0, cfg)
    frames = 70
    x = rng.integers(0, pa.order, size=(frames, n // bps)).astype(np.int64)
    y = pa.constellation[x] + np.sqrt(n0) * rng.normal(size=x.shape)
    for prec in ("fp32", "fp64"):
        for sched in (0, 1):
            for mode in (0, 1, 2):
                out = Reconciler(dec, nm, mode=mode, precision=prec, demap="fast" if prec == "fp32" else "exact",
                                 lanes=32, schedule=sched).run_device(torch.tensor(y, device="cuda"),
                                                                     torch.tensor(x, device="cuda"), 12, k_info=n // 2)
    dec.decode(np.zeros(n), np.zeros(dec.cnum, dtype=np.uint8), 3)
    dec.check_lappr(np.ones(n), np.zeros(dec.cnum, dtype=np.uint8))
    c2v = np.zeros(dec.ednum); v2c = rng.normal(size=dec.ednum)
    dec.process_check_node(0, np.zeros(dec.cnum, dtype=np.uint8), c2v, v2c)
    dec.process_var_node(0, np.zeros(n), c2v, v2c, np.zeros(n))
    torch.cuda.synchronize()
print("sanitize case done")
