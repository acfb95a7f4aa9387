#!/usr/bin/env python3
"""One launch of each kernel of the extended NoiseMapper surface and of the mutual-information Monte Carlo, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import mutual_information as mi

pa = qr.PAMAlphabet(2, 2); n0 = pa.variance * 10 ** (-4.0 / 10) / 2
nm = qr.NoiseMapper(pa, n0, np.array([0, 1, 0, 1], dtype=np.uint8))
g = torch.Generator(device="cuda"); g.manual_seed(1)
n = 256 * 32400
nh = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
x = torch.randint(0, 4, (n,), device="cuda", generator=g)
y = torch.tensor(pa.constellation, device="cuda")[x] + float(np.sqrt(n0)) * torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
for rep in range(2):
    a = nm.demap_noise_batch(nh, x)
    b = nm.demap_lappr_simplified_array_batch(nh, x)
    c = nm.demap_lappr_sofisticated_array_batch(nh, x)
    d = nm.F_Y_batch(y)
    e = mi.information_from_samples(nm, mi.P_xhat(nm), x, y, mode="fast")
    torch.cuda.synchronize()
print("I(X;Xhat), I(X;Y), I(X,N;Xhat) terms:", e)
