#!/usr/bin/env python3
"""Stage timings and decoder sweeps on the GPU (development tool, not part of the bench contract).

    python tools/sweep_decode.py [--n 64800 --frames 1024 --snr 3.0 --lanes 32,64,128 --schedules 0,1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import codes


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64800)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--snr", type=float, default=3.0)
    ap.add_argument("--lanes", default="32,64,96,128,256,512,1024")
    ap.add_argument("--schedules", default="0,2")
    ap.add_argument("--precisions", default="fp32")
    ap.add_argument("--maxiter", type=int, default=50)
    ap.add_argument("--stages", action="store_true")
    ap.add_argument("--fused", default="32:1", help="tile:hints[:pp_lag[:pp_items[:rows_per_claim]]] combinations tried for schedule 2 (fused)")
    a = ap.parse_args()
    n, B = a.n, a.frames
    vid, cid = codes.regular_ldpc(n, 3, 6, seed=1)
    E = vid.size; C = n // 2
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid); pa = qr.PAMAlphabet(2, 2)
    cfg = np.array([0, 1, 0, 1], dtype=np.uint8)
    n0 = pa.variance * 10 ** (-a.snr / 10) / 2
    nm = qr.NoiseMapper(pa, n0, cfg)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    const = torch.tensor(pa.constellation, device="cuda")
    x = torch.randint(0, 4, (B, n // 2), device="cuda", generator=gen)
    y = const[x] + float(np.sqrt(n0)) * torch.randn(x.shape, device="cuda", dtype=torch.float64, generator=gen)
    t_fe, (idx, nh, word) = timeit(lambda: nm.front_end_batch(y))
    t_sy, synd = timeit(lambda: mat.eval_syndrome_batch(word))
    t_df, llr = timeit(lambda: nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float32))
    t_d32, llr32 = timeit(lambda: nm.demap_lappr_array_batch(nh, x, mode="fast32", out_dtype=torch.float32))
    print(f"frames {B}  n {n}  snr {a.snr}")
    print(f"demap(fast32,f32) {t_d32:8.2f} ms   max |fast32-fast| = {(llr32 - llr).abs().max().item():.3e}   "
          f"max rel = {((llr32 - llr).abs() / llr.abs().clamp_min(1e-3)).max().item():.3e}")
    print(f"front_end {t_fe:8.2f} ms   syndrome {t_sy:8.2f} ms   demap(fast,f32) {t_df:8.2f} ms")
    if a.stages:
        t_de, llr_e = timeit(lambda: nm.demap_lappr_array_batch(nh, x, mode="exact", out_dtype=torch.float32), reps=1)
        print(f"demap(exact,f32) {t_de:8.2f} ms   max |fast-exact| = {(llr - llr_e).abs().max().item():.3e}")
    for prec in a.precisions.split(","):
        w = 4 if prec == "fp32" else 8
        bpi = 4 * E * w + 2 * n * w + C
        inp = llr if prec == "fp32" else llr.double()
        for sched in [int(s) for s in a.schedules.split(",")]:
          for combo in (a.fused.split(",") if sched == 2 else [""]):
            if combo:
                parts = (combo.split(":") + ["1", "0", "4"])[:5] if combo.count(":") < 4 else combo.split(":")
                os.environ["QAMRECON_FUSED_TILE"], os.environ["QAMRECON_FUSED_HINTS"] = parts[0], parts[1]
                os.environ["QAMRECON_FUSED_PP_LAG"], os.environ["QAMRECON_FUSED_PP_ITEMS"] = parts[2], parts[3]
                os.environ["QAMRECON_FUSED_RPC"] = parts[4]
            for lanes in [int(s) for s in a.lanes.split(",")]:
                t, (ok, it, post) = timeit(lambda: dec.decode_batch(inp, synd, a.maxiter, precision=prec, lanes=lanes,
                                                                   schedule=sched))
                fi, steps = dec.last_stats(prec, lanes)
                print(f"{prec} sched {sched}{' ' + combo if combo else ''} lanes {lanes:5d}: {t:9.2f} ms  {B / t * 1e3:9.0f} frames/s  "
                      f"{fi * bpi / t / 1e6:8.0f} GB/s alg  avg it {fi / B:5.1f}  steps {steps}  ok {int(ok.sum())}",
                      flush=True)
                # free the workspace of this lane count before the next one
                for key in list(dec._dec):
                    from qamreconciliation import _abi
                    _abi.lib().qr_decoder_destroy(dec._dec.pop(key))


if __name__ == "__main__":
    main()
