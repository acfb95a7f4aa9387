#!/usr/bin/env python3
"""Decoder throughput on the other BASELINE configurations (supplementary to bench.py, which measures config 2):
config 3 (irregular n = 131 070, R = 0.2, 8-PAM, 100 iterations max) and config 4 (QKD scale, n = 2^20,
R = 0.1, 2-PAM), worst case (every frame runs to the iteration limit), fp32, schedule chosen by QR_SCHED_AUTO.

    python tools/config_throughput.py [--frames3 512 --frames4 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch

import qamreconciliation as qr
from qamreconciliation import codes


def run(name, vid, cid, bps, snr, cfg, frames, maxiter, lanes):
    n, c = int(vid.max()) + 1, int(cid.max()) + 1
    E = vid.size
    dec = qr.Decoder(vid, cid); mat = qr.Matrix(vid, cid); pa = qr.PAMAlphabet(bps, 2)
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    nm = qr.NoiseMapper(pa, n0, cfg)
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    x = torch.randint(0, pa.order, (frames, n // bps), device="cuda", generator=gen)
    y = torch.tensor(pa.constellation, device="cuda")[x] + float(np.sqrt(n0)) * torch.randn(
        x.shape, device="cuda", dtype=torch.float64, generator=gen)
    _, nh, word = nm.front_end_batch(y, want_index=False)
    synd = mat.eval_syndrome_batch(word)
    llr = nm.demap_lappr_array_batch(nh, x, mode="fast", out_dtype=torch.float32)
    del y, nh
    for lanes in (lanes if isinstance(lanes, (list, tuple)) else [lanes]):
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ok, it, post = dec.decode_batch(llr, synd, maxiter, precision="fp32", lanes=lanes, schedule=SCHEDULE)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        fi, steps = dec.last_stats("fp32", lanes)
        ms = min(ts[1:])
        bpi = 4 * E * 4 + 2 * n * 4 + c
        print(json.dumps({"config": name, "n": n, "checks": c, "edges": int(E), "frames": frames, "lanes": lanes,
                          "max_iterations": maxiter, "avg_iterations": fi / frames, "converged": int(ok.sum()),
                          "decode_ms": ms, "frames_per_s": frames / ms * 1e3, "edge_updates_per_s": fi * E / ms * 1e3,
                          "algorithmic_GBps": fi * bpi / ms / 1e6}), flush=True)
        from qamreconciliation import _abi
        for key in list(dec._dec):      # free this lane count's workspace before the next one
            _abi.lib().qr_decoder_destroy(dec._dec.pop(key))


SCHEDULE = 3


def main():
    global SCHEDULE
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames3", type=int, default=512)
    ap.add_argument("--frames4", type=int, default=256)
    ap.add_argument("--lanes3", default="256", help="comma list of resident-frame counts tried for config 3")
    ap.add_argument("--lanes4", default="128", help="the same for config 4")
    ap.add_argument("--schedule", type=int, default=3)
    ap.add_argument("--only", type=int, default=0, help="3 or 4: run only that configuration")
    a = ap.parse_args()
    SCHEDULE = a.schedule
    vid, cid = codes.irregular_ldpc(131070, 104856, [3, 8], [0.9, 0.1], seed=3)
    cfg = np.zeros(8, dtype=np.uint8); cfg[1::2] = 1
    if a.only in (0, 3):
        run("3: irregular n=131070 R=0.2 8-PAM", vid, cid, 3, 3.0, cfg, a.frames3, 100, [int(v) for v in a.lanes3.split(",")])
    if a.only == 3:
        return
    vid, cid = codes.irregular_ldpc(1 << 20, 943718, [3, 4, 10], [0.8, 0.15, 0.05], seed=4)
    run("4: irregular n=2^20 R=0.1 2-PAM", vid, cid, 1, -12.0, np.array([0, 1], dtype=np.uint8), a.frames4, 60, [int(v) for v in a.lanes4.split(",")])


if __name__ == "__main__":
    main()
