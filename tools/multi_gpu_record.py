#!/usr/bin/env python3
"""Multi-GPU records that are not bench lines (run under torchrun, one process per GPU):

  1. the host link alone: every rank copies the bench's per-step buffers (pinned f64 samples + i64 symbols in,
     f32 LLRs out) with nothing else running -- aggregate GB/s over all ranks, i.e. what `e2e` can reach at most;
  2. BASELINE config 4 (QKD-scale irregular n = 2^20, R = 0.1, 2-PAM, 60 iterations max) with a batch of frames
     sharded across the ranks, the BER / FER / iteration counters reduced with one all-reduce over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_record.py [--frames4 64]
Prints one JSON line per record on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames4", type=int, default=64, help="config-4 frames per rank and batch")
    ap.add_argument("--batches4", type=int, default=2)
    ap.add_argument("--copy-frames", type=int, default=4096)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- 1. host link alone (the bench's e2e buffers: 4096 frames x 32400 symbols)
    B, S, N = a.copy_frames, 32400, 64800
    hy = torch.empty((B, S), dtype=torch.float64).pin_memory(); hx = torch.empty((B, S), dtype=torch.int64).pin_memory()
    hp = torch.empty((B, N), dtype=torch.float32).pin_memory()
    dy = torch.empty((B, S), dtype=torch.float64, device=dev); dx = torch.empty((B, S), dtype=torch.int64, device=dev)
    dp = torch.empty((B, N), dtype=torch.float32, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    rec = {}
    for name, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                if do_in:
                    with torch.cuda.stream(s_in):
                        dy.copy_(hy, non_blocking=True); dx.copy_(hx, non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s_out):
                        hp.copy_(dp, non_blocking=True)
            barrier()
            el = maxr(time.perf_counter() - t0)
        nbytes = 3 * ((B * S * 16 if do_in else 0) + (B * N * 4 if do_out else 0)) * world
        rec[name + "_GBps_all_ranks"] = nbytes / el / 1e9
    if rank == 0:
        print(json.dumps({"record": "host link alone", "n_gpus": world, **rec,
                          "buffers": "pinned host memory; per rank 2.12 GB in (f64 samples, i64 symbols) and 1.06 GB out (f32 LLRs) per pass"}), flush=True)
    del hy, hx, hp, dy, dx, dp

    # ---- 2. config 4 sharded across the ranks
    import qamreconciliation as qr
    from qamreconciliation import codes
    from qamreconciliation.pipeline import Reconciler
    n, c = 1 << 20, 943718
    vid, cid = codes.irregular_ldpc(n, c, [3, 4, 10], [0.8, 0.15, 0.05], seed=4)
    dec = qr.Decoder(vid, cid); pa = qr.PAMAlphabet(1, 2)
    out_rec = []
    for snr in (-12.0, -4.0):
        n0 = pa.variance * 10 ** (-snr / 10) / 2
        nm = qr.NoiseMapper(pa, n0, np.array([0, 1], dtype=np.uint8))
        rcl = Reconciler(dec, nm, precision="fp32", demap="fast", lanes=a.frames4)
        gen = torch.Generator(device=dev); gen.manual_seed(4000 + rank)
        F = a.frames4
        x = torch.randint(0, 2, (F, n), device=dev, generator=gen)
        y = torch.tensor(pa.constellation, device=dev)[x] + float(np.sqrt(n0)) * torch.randn((F, n), device=dev, dtype=torch.float64, generator=gen)
        rcl.run_device(y, x, 60, k_info=n - c, want_post=False)           # warm-up
        counters = torch.zeros(5, dtype=torch.int64, device=dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.batches4):
            out = rcl.run_device(y, x, 60, k_info=n - c, want_post=False)
            counters += torch.stack([out["bit_errors"].sum(dtype=torch.int64), (out["bit_errors"] > 0).sum(),
                                     out["success"].sum(dtype=torch.int64),
                                     (out["iters"].to(torch.int64) * out["success"].to(torch.int64)).sum(),
                                     torch.tensor(F, dtype=torch.int64, device=dev)])
        if world > 1:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM)
        e1.record()
        barrier()
        ms = maxr(e0.elapsed_time(e1))
        cnt = counters.cpu().tolist()
        frames = F * a.batches4 * world
        out_rec.append({"EsN0_dB": snr, "frames": frames, "frames_per_s": frames / ms * 1e3, "ms": ms,
                        "fer": cnt[1] / cnt[4], "ber": cnt[0] / (cnt[4] * (n - c)), "converged": cnt[2],
                        "avg_iterations_converged": cnt[3] / max(1, cnt[2])})
    if rank == 0:
        print(json.dumps({"record": "BASELINE config 4 sharded", "n_gpus": world, "n": n, "checks": c, "edges": int(vid.size),
                          "frames_per_rank_per_batch": a.frames4, "points": out_rec,
                          "exchange": "one all-reduce of 5 int64 counters per point, inside the timed region"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
