#!/usr/bin/env python3
"""Demapper cost split on the GPU: whole fast demapper vs. one inverse per element (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np, torch
import qamreconciliation as qr
pa = qr.PAMAlphabet(2, 2); n0 = pa.variance * 10 ** (-3.0 / 10) / 2
nm = qr.NoiseMapper(pa, n0, np.array([0, 1, 0, 1], dtype=np.uint8))
g = torch.Generator(device="cuda"); g.manual_seed(1)
n = 2048 * 32400
nh = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
x = torch.randint(0, 4, (n,), device="cuda", generator=g)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
print("demap fast  f32 %.2f ms" % t(lambda: nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float32)))
print("demap fast  f64 %.2f ms" % t(lambda: nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float64)))
for r in range(4):
    reg = torch.full((n,), r, device="cuda", dtype=torch.int64)
    print("g_inv_search fast region %d: %.2f ms" % (r, t(lambda: nm.g_inv_search_batch(nh, reg, mode="fast"))))
print("copy-like (bare_llr f32) %.2f ms" % t(lambda: nm.bare_llr_batch(x, out_dtype=torch.float32)))
