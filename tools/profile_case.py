#!/usr/bin/env python3
"""One small pass of the whole chain (config-2 code, 512 frames, 10 iterations max) for ncu captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import codes
from qamreconciliation.pipeline import Reconciler

frames, maxiter = int(os.environ.get("PROF_FRAMES", "512")), int(os.environ.get("PROF_MAXITER", "10"))
vid, cid = codes.regular_ldpc(64800, 3, 6, seed=1)
dec = qr.Decoder(vid, cid); pa = qr.PAMAlphabet(2, 2)
n0 = pa.variance * 10 ** (-3.0 / 10) / 2
nm = qr.NoiseMapper(pa, n0, np.array([0, 1, 0, 1], dtype=np.uint8))
rec = Reconciler(dec, nm, precision="fp32", demap="fast", schedule=int(os.environ.get("PROF_SCHEDULE", "0")))
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.randint(0, 4, (frames, 32400), device="cuda", generator=gen)
y = torch.tensor(pa.constellation, device="cuda")[x] + float(np.sqrt(n0)) * torch.randn(x.shape, device="cuda", dtype=torch.float64, generator=gen)
for rep in range(2):          # rep 0 warms up, rep 1 is what the profiler keeps (-s skips rep 0's launches)
    out = rec.run_device(y, x, maxiter, k_info=32400)
    torch.cuda.synchronize()
print("iters", out["iters"][:4].tolist(), "errors", out["bit_errors"][:4].tolist())
