#!/usr/bin/env python3
"""Per-stage CUDA-event timing of one bench step (4096 frames of config 2), to see what the step is made of."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import _abi, codes, utils
from qamreconciliation._util import stream

B = int(os.environ.get("FRAMES", "4096"))
vid, cid = codes.regular_ldpc(64800, 3, 6, seed=1)
dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid); pa = qr.PAMAlphabet(2, 2)
n0 = pa.variance * 10 ** (-3.0 / 10) / 2
nm = qr.NoiseMapper(pa, n0, np.array([0, 1, 0, 1], dtype=np.uint8))
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
x = torch.randint(0, 4, (B, 32400), device="cuda", generator=gen)
y = torch.tensor(pa.constellation, device="cuda")[x] + float(np.sqrt(n0)) * torch.randn(x.shape, device="cuda", dtype=torch.float64, generator=gen)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


for rep in range(3):
    t = [ev()]
    _, nh, word = nm.front_end_batch(y, want_index=False); t.append(ev())
    synd = mat.eval_syndrome_batch(word); t.append(ev())
    llr = nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float32); t.append(ev())
    ok, it, post = dec.decode_batch(llr, synd, 50, precision="fp32", schedule=int(os.environ.get("SCHEDULE", "2")),
                                    lanes=int(os.environ.get("LANES", "1024"))); t.append(ev())
    err = utils.count_errors_batch(post, word, k=32400); t.append(ev())
    torch.cuda.synchronize()
    names = ["front_end", "syndrome", "demap", "decode", "count_errors"]
    print("rep", rep, "  ".join(f"{n} {t[i].elapsed_time(t[i + 1]):7.2f}" for i, n in enumerate(names)),
          f" total {t[0].elapsed_time(t[-1]):7.2f} ms")
