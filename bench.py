#!/usr/bin/env python3
"""bench.py -- headline benchmark of the reverse-reconciliation hot path (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic frames per GPU:
    y, x  ->  hard decision + softening metric + Gray bits  ->  syndrome  ->  Alice's LLRs
          ->  syndrome sum-product decoding (50 iterations max)  ->  bit-error count
on the configuration the metric is quoted on (BASELINE config 2: synthetic (3,6)-regular LDPC,
n = 64 800, R = 1/2, 4-PAM, Alternating configuration, batch of 4096 frames per GPU), at an Es/N0
below the waterfall so that every frame runs all 50 iterations (the worst case and the headline).

    python bench.py [--gpus N --steps K --warmup W]        # our arm (torchrun for N > 1)
    python bench.py --impl reference [...]                 # the reference's CPU path on the host cores

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same
metric through the host-buffer C-ABI call (qr_reconcile_host) with the copies inside the timed
region; `roofline` = the decoder kernel against the measured HBM peak; `cpu_baseline` = the CPU
oracle port timed on this box's cores on a bounded sample of the same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "qam-reconciliation_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_CODE, DV, DC, CODE_SEED = 64800, 3, 6, 1
BPS = 2
MAXITER = 50
METRIC = "decoded frames/s (n=64800 R=1/2 4-PAM soft reverse reconciliation, 50 iterations max)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--snr", type=float, default=3.0, help="Es/N0 [dB]; 3.0 is below the waterfall")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--demap", default="fast", choices=["fast", "exact"])
    ap.add_argument("--lanes", type=int, default=-1,
                    help="frames resident in the decoder (0 = library default; -1 = 1024 for the fused schedule, else 0)")
    ap.add_argument("--schedule", type=int, default=-1,
                    help="0 persistent two-phase kernel, 1 launch per phase, 2 fused flooding iteration (persistent); "
                         "-1 = 2")
    ap.add_argument("--n", type=int, default=N_CODE, help="code length (default: the metric's 64800)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per CPU worker (0 = auto)")
    a = ap.parse_args()
    if a.schedule < 0:
        a.schedule = 2
    if a.lanes < 0:
        a.lanes = (1024 if a.precision == "fp32" else 512) if a.schedule == 2 else 0
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """Block until nvidia-smi has delivered its first sample: its start-up (NVML initialisation holds driver
        locks for ~100 ms and delays kernel launches) must not fall into the timed region."""
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)
        self.skip = len(self.samples)     # samples taken before the timed region are not "under load"

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        skip = getattr(self, "skip", 0)
        for s in (self.samples[skip:] or self.samples):
            f = [t.strip() for t in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
_W = {}


def _cpu_init(kind, n, snr):
    """Per worker process, once: build the code, the decoder (the reference's constructor is quadratic:
    ~40-100 s at n = 64800) and the mapper.  kind: 'port' (oracle C restatement) or 'reference'
    (compiled reference in oracle/_ref)."""
    import importlib.util
    t_setup = time.time()
    spec = importlib.util.spec_from_file_location("qr_codes", os.path.join(PKG, "qamreconciliation", "codes.py"))
    codes = importlib.util.module_from_spec(spec)     # the product's numpy-only code generator
    spec.loader.exec_module(codes)
    vid, cid = codes.regular_ldpc(n, DV, DC, seed=CODE_SEED)
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        for k in [k for k in sys.modules if k == "qamreconciliation" or k.startswith("qamreconciliation.")]:
            del sys.modules[k]
        import qamreconciliation as ref
        pa = ref.PAMAlphabet(BPS, 2)
        dec = ref.Decoder(vid.copy(), cid.copy()); mat = ref.Matrix(vid.copy(), cid.copy())
        NM = ref.NoiseMapper
    else:
        from oracle import port as orc
        pa = orc.PAMAlphabet(BPS, 2)
        dec = orc.Decoder(vid, cid); mat = orc.Matrix(vid, cid)
        NM = orc.NoiseMapper
    cfg = np.zeros(1 << BPS, dtype=np.uint8); cfg[1::2] = 1
    n0 = pa.variance * 10 ** (-snr / 10) / 2
    _W.update(kind=kind, n=n, n0=n0, cfg=cfg, pa=pa, dec=dec, mat=mat, nm=NM(pa, n0, cfg),
              a=np.asarray(pa.constellation), setup=time.time() - t_setup, onm=None)


def _cpu_step(args):
    """One frame at a time through the whole chain (the north star's CPU protocol)."""
    seeds, demap_div = args
    W = _W
    n, n0, pa, nm, dec, mat, a = W["n"], W["n0"], W["pa"], W["nm"], W["dec"], W["mat"], W["a"]
    S = n // BPS
    t_chain = t_demap = t_dec = 0.0
    iters = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        x = rng.integers(0, 1 << BPS, size=S).astype(np.int64)
        y = a[x] + np.sqrt(n0) * rng.normal(size=S)
        t0 = time.time()
        xh = np.array(nm.hard_decide_index(y.copy()), dtype=np.int64)
        nh = np.array(nm.map_noise(y.copy(), xh.copy()))
        word = np.array(pa.demap_symbols_to_bits(xh.copy()))
        if word.dtype.kind == "S":
            word = word.view(np.uint8)
        synd = np.array(mat.eval_syndrome(word.copy()), dtype=np.uint8)
        t1 = time.time()
        sub = S // demap_div
        lap_part = np.array(nm.demap_lappr_array(nh[:sub].copy(), x[:sub].copy()))
        t2 = time.time()
        if demap_div > 1:
            # the decoder needs LLRs for the whole frame: the oracle port supplies the rest (untimed)
            if W["onm"] is None:
                from oracle import port as orc2
                W["onm"] = orc2.NoiseMapper(orc2.PAMAlphabet(BPS, 2), n0, W["cfg"])
            lap = np.concatenate([lap_part, W["onm"].demap_lappr_array(nh[sub:], x[sub:])])
        else:
            lap = lap_part
        t3 = time.time()
        ok, it, post = dec.decode(lap.copy(), synd.copy(), MAXITER)
        t4 = time.time()
        t_chain += t1 - t0; t_demap += (t2 - t1) * demap_div; t_dec += t4 - t3
        iters += int(it)
    return dict(setup=W["setup"], chain=t_chain, demap=t_demap, dec=t_dec, frames=len(seeds), iters=iters)


class CpuArm:
    """A pool of worker processes, each with its own decoder, reused across steps."""

    def __init__(self, kind, n, snr, workers=None):
        import multiprocessing as mp
        self.P = workers or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.P, initializer=_cpu_init, initargs=(kind, n, snr))
        self.step_no = 0

    def step(self, frames_per_worker, demap_div):
        jobs = [([100000 * self.step_no + 1000 * w + f for f in range(frames_per_worker)], demap_div)
                for w in range(self.P)]
        self.step_no += 1
        t0 = time.time()
        res = self.pool.map(_cpu_step, jobs, chunksize=1)
        wall = time.time() - t0
        per_frame = [(r["chain"] + r["demap"] + r["dec"]) / r["frames"] for r in res]
        # every worker runs concurrently on its own core: aggregate rate = sum of per-worker rates
        return dict(fps=sum(1.0 / t for t in per_frame), cores=self.P, wall=wall,
                    setup=max(r["setup"] for r in res),
                    chain=float(np.mean([r["chain"] / r["frames"] for r in res])),
                    demap=float(np.mean([r["demap"] / r["frames"] for r in res])),
                    dec=float(np.mean([r["dec"] / r["frames"] for r in res])),
                    iters=float(np.mean([r["iters"] / r["frames"] for r in res])))

    def close(self):
        self.pool.close()
        self.pool.join()


def run_cpu(kind, n, snr, frames_per_worker, demap_div, workers=None):
    arm = CpuArm(kind, n, snr, workers)
    try:
        return arm.step(frames_per_worker, demap_div)
    finally:
        arm.close()


def reference_arm(a):
    """The reference's own CPU implementation of the path on this box's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build_ref
    kind = "reference" if build_ref.have_ref() else "port"
    # The reference's demap_lappr_array costs ~27 s per n=64800 frame (Python-level scipy.erf per call):
    # it is timed on the first 1/16 of each frame's symbols and scaled by 16, everything else on whole frames.
    div = 16 if kind == "reference" else 1
    arm = CpuArm(kind, a.n, a.snr)
    results = [arm.step(1, div) for _ in range(a.warmup + a.steps)]
    arm.close()
    res = results[a.warmup:] or results
    fps = float(np.mean([r["fps"] for r in res]))
    ms = float(np.mean([1000.0 * r["cores"] / r["fps"] for r in res]))
    K = a.n // 2
    sample = (f"per step: {res[0]['cores']} worker processes x 1 frame each, one frame per process; "
              f"hard decision/map_noise/bits/syndrome/decode on the whole frame, demap_lappr_array on "
              f"1/{div} of the symbols scaled x{div}; decoder constructor ({res[0]['setup']:.0f} s) excluded; "
              f"per frame: chain {res[0]['chain']:.3f} s, demap {res[0]['demap']:.2f} s, decode {res[0]['dec']:.2f} s")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.n, a.snr), "frames_per_step": res[0]["cores"]},
            "info_gbit_per_s": fps * K / 1e9,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": res[0]["cores"], "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
def ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build_library()
    if world > 1:
        dist.barrier()
    import qamreconciliation as qr
    from qamreconciliation import codes
    from qamreconciliation.pipeline import Reconciler

    n = a.n
    vid, cid = codes.regular_ldpc(n, DV, DC, seed=CODE_SEED)
    E, C = vid.size, n * DV // DC
    K = n - C
    S = n // BPS
    dec = qr.Decoder(vid, cid)
    pa = qr.PAMAlphabet(BPS, 2)
    cfg = np.zeros(pa.order, dtype=np.uint8); cfg[1::2] = 1
    n0 = pa.variance * 10 ** (-a.snr / 10) / 2
    nm = qr.NoiseMapper(pa, n0, cfg)
    rec = Reconciler(dec, nm, precision=a.precision, demap=a.demap, lanes=a.lanes or None, schedule=a.schedule)
    B = a.frames
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    const = torch.tensor(pa.constellation, device=dev)
    # distinct inputs per step (working set >> L2: 4096 x 32400 x 16 B = 2.1 GB per step)
    n_sets = 2
    xs = [torch.randint(0, pa.order, (B, S), device=dev, generator=gen) for _ in range(n_sets)]
    ys = [const[x] + float(np.sqrt(n0)) * torch.randn((B, S), device=dev, dtype=torch.float64, generator=gen)
          for x in xs]
    w = 4 if a.precision == "fp32" else 8
    bytes_per_frame_iter = 4 * E * w + 2 * n * w + C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        return rec.run_device(ys[i % n_sets], xs[i % n_sets], MAXITER, k_info=K)

    # ---- device-resident leg
    # nvidia-smi is started BEFORE the warm-up: its start-up takes ~0.2 s, and a GPU left idle that long drops its
    # clocks and spends the first ~25 ms of the timed region ramping them up again
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for i in range(a.warmup):
        out = step(i)
    barrier()
    sampler.skip = len(sampler.samples)       # samples taken before the timed region do not count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dec_ev = []
    counters = torch.zeros(5, dtype=torch.int64, device=dev)
    b_frames = torch.tensor(B, dtype=torch.int64, device=dev)
    barrier()
    ev[0].record()
    frame_iters = 0
    for i in range(a.steps):
        out = step(a.warmup + i)
        # (b_frames lives on the device: building it here would be a synchronous host-to-device copy every step,
        # which stalls the launch queue behind the whole step)
        counters += torch.stack([out["bit_errors"].sum(dtype=torch.int64), (out["bit_errors"] > 0).sum(),
                                 out["success"].sum(dtype=torch.int64),
                                 (out["iters"].to(torch.int64) * out["success"].to(torch.int64)).sum(),
                                 b_frames])
    ev[1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)     # the only data exchange of the path
    elapsed_ms = float(t.item())
    total_frames = B * a.steps * world
    value = total_frames / (elapsed_ms / 1e3)

    # ---- decoder kernel alone (roofline): same inputs, events around the decode call only
    out_s = rec.run_device(ys[0], xs[0], MAXITER, k_info=K, stagewise=True)
    llr = out_s["llr"]; synd = out_s["synd"]
    del out_s
    for _ in range(2):
        dec.decode_batch(llr, synd, MAXITER, precision=a.precision, lanes=a.lanes or None, schedule=a.schedule)
    torch.cuda.synchronize()
    kt = []
    for _ in range(max(3, a.steps)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ok_, it_, _ = dec.decode_batch(llr, synd, MAXITER, precision=a.precision, lanes=a.lanes or None,
                                       schedule=a.schedule)
        e1.record()
        torch.cuda.synchronize()
        kt.append(e0.elapsed_time(e1))
    fi, steps_exec = dec.last_stats(a.precision, a.lanes or None)
    k_ms = float(np.mean(kt))
    peak, peak_src = peaks()
    achieved = fi * bytes_per_frame_iter / (k_ms / 1e3) / 1e9
    traffic = None
    try:   # DRAM bytes of the same launch from the committed ncu --set full captures (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        for ent in tj.get("entries", [tj]):
            c = ent["config"]
            sched = {"persistent": 0, "launch": 1, "fused": 2}[c["schedule"]]
            if (c["n"], c["frames"], c["max_iterations"], c["precision"], sched, c["lanes"]) == \
                    (n, B, MAXITER, a.precision, a.schedule, a.lanes or 512) and fi == B * MAXITER:
                traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    kname = {0: "k_persistent (decoder, one launch per batch)", 1: "k_check+k_var",
             2: "k_fused (decoder, fused flooding iteration, one launch per batch)"}[a.schedule]
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": fi * bytes_per_frame_iter, "launch_ms": k_ms,
                "edge_updates_per_s": fi * E / (k_ms / 1e3), "frame_iterations_per_launch": fi,
                "decode_only_frames_per_s": B / (k_ms / 1e3)}

    # ---- end to end through the host-buffer C ABI
    e2e = None
    if not a.no_e2e:
        hy = [y.cpu().pin_memory() for y in ys]
        hx = [x.cpu().pin_memory() for x in xs]
        outs = dict(success=torch.empty(B, dtype=torch.uint8).pin_memory(),
                    iters=torch.empty(B, dtype=torch.int32).pin_memory(),
                    bit_errors=torch.empty(B, dtype=torch.int32).pin_memory(),
                    post=torch.empty((B, n), dtype=torch.float32 if a.precision == "fp32" else torch.float64).pin_memory())
        for i in range(min(a.warmup, 2)):
            rec.run_host(hy[i % n_sets], hx[i % n_sets], MAXITER, K, outs)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            rec.run_host(hy[i % n_sets], hx[i % n_sets], MAXITER, K, outs)
        barrier()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = B * S * 16
        d2h = B * (1 + 4 + 4) + outs["post"].numel() * outs["post"].element_size()
        e2e = {"value": total_frames / float(te.item()), "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "api": "qr_reconcile_host: host y (f64) + tx symbols (i64) in; success, iterations, bit errors and "
                      "final LLRs out; pinned host memory"}

    if rank != 0:
        return
    cpu = None
    if not a.no_cpu:
        fpw = a.cpu_frames or (2 if n >= 30000 else 8)
        r = run_cpu("port", n, a.snr, fpw, 1)
        cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['cores']} processes x {fpw} frames of the same workload, one frame per process at a "
                         f"time (oracle/qr_oracle.c, gcc -O2); per frame: front end+syndrome {r['chain']:.3f} s, "
                         f"demap {r['demap']:.2f} s, decode {r['dec']:.2f} s ({r['iters']:.0f} iterations)"}
    cnt = counters.cpu().tolist()
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, a.snr),
                       "frames_per_gpu_per_step": B, "demap": a.demap, "decoder_lanes": a.lanes or 512,
                       "schedule": {0: "persistent", 1: "launch", 2: "fused"}[a.schedule],
                       "l2_policy": f"{n_sets} alternating input sets of {B * S * 16 / 1e9:.1f} GB each (>> 126 MB L2)"},
            "info_gbit_per_s": value * K / 1e9,
            "avg_iterations": fi / B, "ber": cnt[0] / max(1, cnt[4] * K), "fer": cnt[1] / max(1, cnt[4]),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            # per step: front end, syndrome, demapper, batch init, persistent decoder, error count
            "gpu_launches": 6 * a.steps * world, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def workload_name(n, snr):
    return (f"(3,6)-regular LDPC n={n} R=1/2 (seed {CODE_SEED}), 4-PAM Alternating, Es/N0={snr} dB, "
            f"maxiter {MAXITER}, soft reverse reconciliation (BASELINE config 2)")


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
