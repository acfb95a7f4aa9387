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
oracle port timed on this box's cores on frames taken from the GPU batch, with its agreement with the
GPU's fp64 mode on those frames.  At N = 1 the line also carries `operating_points`: the same measurement
in the waterfall (4 dB) and above it (5 dB), the fp64 parity mode, and the decoder on BASELINE configs 3 and
4 (`--no-extras` skips them).
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
                    help="frames resident in the decoder (0 = the package's default: one lane per frame up to 4096; -1 = one lane "
                         "per frame of the step for the fused fp32 schedule (4096), 512 for fp64, else 0)")
    ap.add_argument("--schedule", type=int, default=-1,
                    help="0 persistent two-phase kernel, 1 launch per phase, 2 fused flooding iteration (persistent); "
                         "-1 = 2")
    ap.add_argument("--n", type=int, default=N_CODE, help="code length (default: the metric's 64800)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames per CPU worker (0 = auto)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the supplementary operating points (4 dB, 5 dB, fp64, configs 3 and 4)")
    a = ap.parse_args()
    if a.schedule < 0:
        a.schedule = 2
    if a.lanes < 0:
        # one lane per frame: no lane is ever refilled inside a step (refill generations measured ~8 % slower at 3 dB,
        # ~15 % at 4 dB and ~20 % at 5 dB; the workspace is 1.8 MB per lane of config 2 -- 7.4 GB of the 180 GB)
        a.lanes = ((a.frames + 31) // 32 * 32 if a.precision == "fp32" else 512) if a.schedule == 2 else 0
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
        sm, mx, pw, reasons = [], [], [], set()
        skip = getattr(self, "skip", 0)
        for s in (self.samples[skip:] or self.samples):
            f = [t.strip() for t in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


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
              a=np.asarray(pa.constellation), setup=time.time() - t_setup)


def _cpu_step(args):
    """One WHOLE frame at a time through the whole chain (the north star's CPU protocol): every stage on every
    symbol, nothing scaled.  args = list of (seed | (y, x)) frames: a seed draws the frame here, a (y, x) pair is a
    frame copied out of the GPU batch."""
    W = _W
    n, n0, pa, nm, dec, mat, a = W["n"], W["n0"], W["pa"], W["nm"], W["dec"], W["mat"], W["a"]
    S = n // BPS
    t_chain = t_demap = t_dec = 0.0
    res = []
    for fr in args:
        if isinstance(fr, tuple):
            y, x = np.array(fr[0], dtype=np.float64), np.array(fr[1], dtype=np.int64)
        else:
            rng = np.random.default_rng(fr)
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
        lap = np.array(nm.demap_lappr_array(nh.copy(), x.copy()))
        t2 = time.time()
        ok, it, post = dec.decode(lap.copy(), synd.copy(), MAXITER)
        t3 = time.time()
        t_chain += t1 - t0; t_demap += t2 - t1; t_dec += t3 - t2
        res.append((int(ok), int(it), np.array(post, dtype=np.float64) if isinstance(fr, tuple) else None))
    return dict(setup=W["setup"], chain=t_chain, demap=t_demap, dec=t_dec, frames=len(args),
                iters=sum(r[1] for r in res), results=res)


class CpuArm:
    """A pool of worker processes, each with its own decoder, reused across steps."""

    def __init__(self, kind, n, snr, workers=None):
        import multiprocessing as mp
        self.P = workers or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.P, initializer=_cpu_init, initargs=(kind, n, snr))
        self.step_no = 0
        self.pool.map(_noop, range(self.P), chunksize=1)      # constructors done before anything is timed

    def step(self, frames_per_worker=1, given=None):
        """One step: every worker takes `frames_per_worker` frames, one frame per call; the wall time of the step is
        measured here, around the whole map.  given: list (per worker) of lists of (y, x) frames."""
        if given is not None:
            jobs = given
        else:
            jobs = [[100000 * self.step_no + 1000 * w + f for f in range(frames_per_worker)] for w in range(self.P)]
        self.step_no += 1
        t0 = time.time()
        res = self.pool.map(_cpu_step, jobs, chunksize=1)
        wall = time.time() - t0
        frames = sum(r["frames"] for r in res)
        return dict(fps=frames / wall, frames=frames, cores=self.P, wall=wall,
                    setup=max(r["setup"] for r in res),
                    chain=float(np.mean([r["chain"] / r["frames"] for r in res])),
                    demap=float(np.mean([r["demap"] / r["frames"] for r in res])),
                    dec=float(np.mean([r["dec"] / r["frames"] for r in res])),
                    iters=float(np.mean([r["iters"] / r["frames"] for r in res])),
                    results=[r["results"] for r in res])

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(_):
    time.sleep(0.05)
    return 0


def workload_config(a, n_sets=2):
    """The `config` object of the JSON line: the workload both arms are quoted on."""
    B, S = a.frames, a.n // BPS
    return {"workload": workload_name(a.n, a.snr), "frames_per_gpu_per_step": B,
            "demap": a.demap + (" (fp32 grade: QR_DEMAP_FAST | QR_DEMAP_F32GRADE)" if a.demap == "fast" and a.precision == "fp32" else ""),
            "decoder_lanes": a.lanes or auto_lanes(a.frames), "schedule": {0: "persistent", 1: "launch", 2: "fused", 3: "auto"}[a.schedule],
            "l2_policy": f"{n_sets} alternating input sets of {B * S * 16 / 1e9:.1f} GB each (>> 126 MB L2)"}


def reference_arm(a):
    """The reference's own CPU implementation of the path on this box's cores: the UNMODIFIED compiled reference
    (oracle/_ref) where it was built, else the oracle port.  One step = every host core decodes ONE whole frame of the
    workload (all stages on all symbols, nothing scaled or modelled); `ms_per_step` is the measured wall time of a
    step and `value` = frames of the step / that wall time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build_ref
    kind = "reference" if build_ref.have_ref() else "port"
    arm = CpuArm(kind, a.n, a.snr)
    results = [arm.step(1) for _ in range(a.warmup + a.steps)]
    arm.close()
    res = results[a.warmup:] or results
    wall = float(np.sum([r["wall"] for r in res]))
    frames = int(np.sum([r["frames"] for r in res]))
    fps = frames / wall
    ms = 1000.0 * wall / len(res)
    K = a.n // 2
    r0 = res[0]
    sample = (f"per step: {r0['cores']} worker processes x 1 whole frame each (one frame per process and call), every "
              f"stage on all {a.n // BPS} symbols; wall time of the step measured around the pool; decoder constructor "
              f"({r0['setup']:.0f} s per process) outside the steps; per frame: hard decision + map_noise + bits + syndrome "
              f"{r0['chain']:.3f} s, demap_lappr_array {r0['demap']:.2f} s, decode {r0['dec']:.2f} s "
              f"({r0['iters']:.0f} iterations)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "frames_per_step": r0["frames"],
            "info_gbit_per_s": fps * K / 1e9,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": r0["cores"], "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
class Workload:
    """Code, decoder, mapper and synthetic inputs of one operating point on this rank's GPU."""

    def __init__(self, torch, qr, codes, dev, rank, vid, cid, bps, snr, cfg, frames, maxiter, precision, demap, lanes,
                 schedule, n_sets=2, dec=None):
        from qamreconciliation.pipeline import Reconciler
        self.torch, self.dev = torch, dev
        self.n, self.C, self.E = int(vid.max()) + 1, int(cid.max()) + 1, int(vid.size)
        self.K, self.S, self.bps = self.n - self.C, self.n // bps, bps
        self.B, self.maxiter, self.precision, self.lanes, self.schedule = frames, maxiter, precision, lanes, schedule
        self.dec = dec if dec is not None else qr.Decoder(vid, cid)
        self.pa = qr.PAMAlphabet(bps, 2)
        self.n0 = self.pa.variance * 10 ** (-snr / 10) / 2
        self.nm = qr.NoiseMapper(self.pa, self.n0, cfg)
        self.rec = Reconciler(self.dec, self.nm, precision=precision, demap=demap, lanes=lanes or None, schedule=schedule)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        const = torch.tensor(self.pa.constellation, device=dev)
        # distinct inputs per step (config 2: 4096 x 32400 x 16 B = 2.1 GB per set, >> L2)
        self.n_sets = n_sets
        self.xs = [torch.randint(0, self.pa.order, (frames, self.S), device=dev, generator=gen) for _ in range(n_sets)]
        self.ys = [const[x] + float(np.sqrt(self.n0)) * torch.randn((frames, self.S), device=dev, dtype=torch.float64,
                                                                    generator=gen) for x in self.xs]
        self.w = 4 if precision == "fp32" else 8
        self.bytes_per_frame_iter = 4 * self.E * self.w + 2 * self.n * self.w + self.C

    def step(self, i):
        return self.rec.run_device(self.ys[i % self.n_sets], self.xs[i % self.n_sets], self.maxiter, k_info=self.K)

    def device_leg(self, steps, warmup, barrier, all_reduce=None, sampler=None):
        """W untimed + K timed steps with the inputs resident in HBM; the counters (and, on N > 1, their all-reduce:
        the path's only exchange) are inside the timed region.  Returns (elapsed ms on this rank, counters)."""
        torch = self.torch
        # (b_frames lives on the device: building it per step would be a synchronous host-to-device copy every step,
        # which stalls the launch queue behind the whole step)
        b_frames = torch.tensor(self.B, dtype=torch.int64, device=self.dev)

        def tally(counters, out):
            counters += torch.stack([out["bit_errors"].sum(dtype=torch.int64), (out["bit_errors"] > 0).sum(),
                                     out["success"].sum(dtype=torch.int64),
                                     (out["iters"].to(torch.int64) * out["success"].to(torch.int64)).sum(),
                                     b_frames])

        # The warm-up steps do EVERYTHING a timed step does, in the same way.  Two things used to fall into the timed
        # region otherwise, both between its first and second step (per-step event marks, QAMRECON_BENCH_STEP_TIMES):
        # the first need for a SECOND set of output buffers (the previous step's outputs are still referenced while
        # the next step allocates: a cudaMalloc of ~1.4 GB, ~100 ms), and the first use of torch's reduction kernels
        # in the tally (lazy module loading waits for the device: 25-45 ms).
        out, scratch = None, torch.zeros(5, dtype=torch.int64, device=self.dev)
        for i in range(warmup):
            out = self.step(i)
            tally(scratch, out)
        if all_reduce is not None:
            all_reduce(scratch)
        barrier()
        if sampler is not None:
            sampler.skip = len(sampler.samples)       # samples taken before the timed region do not count
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        counters = torch.zeros(5, dtype=torch.int64, device=self.dev)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)] if os.environ.get("QAMRECON_BENCH_STEP_TIMES") else []
        for m in marks:
            m.record()                                # (created here, outside the timed region)
        barrier()
        ev[0].record()
        for i in range(steps):
            out = self.step(warmup + i)
            if marks:                                 # diagnostic: per-step device times on stderr
                marks[i].record()
            tally(counters, out)
        if all_reduce is not None:
            all_reduce(counters)                      # BER / FER / iteration counters: one tiny NCCL all-reduce
        ev[1].record()
        barrier()
        if marks:
            torch.cuda.synchronize()
            t = [ev[0].elapsed_time(m) for m in marks]
            sys.stderr.write("step end times [ms]: " + " ".join(f"{x:.1f}" for x in t) + "\n")
        return ev[0].elapsed_time(ev[1]), counters

    def decoder_leg(self, reps):
        """The decoder kernel alone on the LLRs / syndromes of input set 0: CUDA events around the launch (torch's
        current stream is the stream the library launches on).  Returns (mean ms, frame-iterations, steps)."""
        torch = self.torch
        out_s = self.rec.run_device(self.ys[0], self.xs[0], self.maxiter, k_info=self.K, stagewise=True)
        llr, synd = out_s["llr"], out_s["synd"]
        del out_s
        kw = dict(precision=self.precision, lanes=self.lanes or None, schedule=self.schedule)
        for _ in range(2):
            self.dec.decode_batch(llr, synd, self.maxiter, **kw)
        torch.cuda.synchronize()
        kt = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.dec.decode_batch(llr, synd, self.maxiter, **kw)
            e1.record()
            torch.cuda.synchronize()
            kt.append(e0.elapsed_time(e1))
        fi, steps_exec = self.dec.last_stats(self.precision, self.lanes or None)
        return float(np.mean(kt)), fi, steps_exec

    def e2e_leg(self, steps, warmup, barrier, sampler=None):
        """Through qr_reconcile_host: pinned host y / x in, success, iterations, bit errors and final LLRs out, the
        copies inside the timed region.  Returns (seconds on this rank, h2d bytes, d2h bytes per step)."""
        torch = self.torch
        hy = [y.cpu().pin_memory() for y in self.ys]
        hx = [x.cpu().pin_memory() for x in self.xs]
        outs = dict(success=torch.empty(self.B, dtype=torch.uint8).pin_memory(),
                    iters=torch.empty(self.B, dtype=torch.int32).pin_memory(),
                    bit_errors=torch.empty(self.B, dtype=torch.int32).pin_memory(),
                    post=torch.empty((self.B, self.n), dtype=torch.float32 if self.precision == "fp32" else torch.float64).pin_memory())
        for i in range(min(warmup, 2)):
            self.rec.run_host(hy[i % self.n_sets], hx[i % self.n_sets], self.maxiter, self.K, outs)
        barrier()
        if sampler is not None:
            sampler.skip = len(sampler.samples)       # (pinning the host buffers above took seconds of idle GPU)
        t0 = time.perf_counter()
        for i in range(steps):
            self.rec.run_host(hy[i % self.n_sets], hx[i % self.n_sets], self.maxiter, self.K, outs)
        barrier()
        el = time.perf_counter() - t0
        h2d = self.B * self.S * 16
        d2h = self.B * (1 + 4 + 4) + outs["post"].numel() * outs["post"].element_size()
        return el, h2d, d2h

    def e2e_compact_leg(self, steps, warmup, barrier):
        """The same pass over the compact wire format (qr_reconcile_host_compact): float32 samples and byte symbols in,
        hard decisions packed 8 per byte out.  NOT the reference's types: reported under its own key."""
        torch = self.torch
        hy = [y.float().cpu().pin_memory() for y in self.ys]
        hx = [x.to(torch.uint8).cpu().pin_memory() for x in self.xs]
        outs = dict(success=torch.empty(self.B, dtype=torch.uint8).pin_memory(),
                    iters=torch.empty(self.B, dtype=torch.int32).pin_memory(),
                    bit_errors=torch.empty(self.B, dtype=torch.int32).pin_memory(),
                    decisions=torch.empty((self.B, (self.n + 7) // 8), dtype=torch.uint8).pin_memory())
        for i in range(min(warmup, 2)):
            self.rec.run_host_compact(hy[i % self.n_sets], hx[i % self.n_sets], self.maxiter, self.K, outs)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            self.rec.run_host_compact(hy[i % self.n_sets], hx[i % self.n_sets], self.maxiter, self.K, outs)
        barrier()
        el = time.perf_counter() - t0
        return el, self.B * self.S * 5, self.B * (1 + 4 + 4) + outs["decisions"].numel()

    def roofline(self, k_ms, fi, kname, ent=None):
        peak, peak_src = peaks()
        traffic = ent["dram_bytes_per_launch"] if ent else None
        achieved = fi * self.bytes_per_frame_iter / (k_ms / 1e3) / 1e9
        r = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": fi * self.bytes_per_frame_iter, "launch_ms": k_ms,
             "edge_updates_per_s": fi * self.E / (k_ms / 1e3), "frame_iterations_per_launch": fi,
             "decode_only_frames_per_s": self.B / (k_ms / 1e3)}
        if traffic:
            # what the HBM really carried (ncu dram__bytes of the same launch) over the live launch time
            r["dram_GBps"] = traffic / (k_ms / 1e3) / 1e9
            r["dram_frac"] = r["dram_GBps"] / peak
            if ent.get("l2_bytes_per_launch"):
                # what the SMs moved through L2 (lts__t_sectors_srcunit_tex of the same capture, partition-fabric crossings
                # not included) against the builder-measured L2 gather bandwidth (tools/membench.cu, profiles/r2_membench.txt)
                r["l2_GBps"] = ent["l2_bytes_per_launch"] / (k_ms / 1e3) / 1e9
                r["l2_peak_GBps"] = ent.get("l2_peak_GBps")
                r["l2_frac"] = r["l2_GBps"] / ent["l2_peak_GBps"] if ent.get("l2_peak_GBps") else None
            r["traffic_source"] = ent.get("source")
        return r


def auto_lanes(frames):
    """What the package picks when no lane count is named (Decoder._auto_lanes, memory permitting)."""
    want = 512
    while want < min(frames, 4096):
        want *= 2
    return want


def lookup_traffic(n, B, maxiter, precision, schedule, lanes, fi):
    """DRAM bytes of the same launch from the committed ncu --set full captures (profiles/r*_traffic.json)."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        try:
            tj = json.load(open(path))
        except Exception:
            continue
        for ent in tj.get("entries", []):
            c = ent["config"]
            sched = {"persistent": 0, "launch": 1, "fused": 2}[c["schedule"]]
            if (c["n"], c["frames"], c["max_iterations"], c["precision"], sched, c["lanes"]) == \
                    (n, B, maxiter, precision, schedule, lanes) and fi == ent.get("frame_iterations", B * maxiter):
                best = ent                                 # later rounds override earlier ones
    return best


KERNEL_NAMES = {0: "k_persistent (decoder, two-phase, one launch per batch)", 1: "k_check+k_var",
                2: "k_fused (decoder, fused flooding iteration, one launch per batch)",
                3: "decoder (schedule chosen by the library)"}


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

    dev = torch.device("cuda", local)
    n = a.n
    vid, cid = codes.regular_ldpc(n, DV, DC, seed=CODE_SEED)
    cfg = np.zeros(1 << BPS, dtype=np.uint8); cfg[1::2] = 1
    wl = Workload(torch, qr, codes, dev, rank, vid, cid, BPS, a.snr, cfg, a.frames, MAXITER, a.precision, a.demap,
                  a.lanes, a.schedule)
    B, K, S = wl.B, wl.K, wl.S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_reduce_sum(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg
    # nvidia-smi is started BEFORE the warm-up: its start-up takes ~0.2 s, and a GPU left idle that long drops its
    # clocks and spends the first ~25 ms of the timed region ramping them up again
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    elapsed_ms, counters = wl.device_leg(a.steps, a.warmup, barrier, all_reduce_sum, sampler)
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = max_over_ranks(elapsed_ms)
    total_frames = B * a.steps * world
    value = total_frames / (elapsed_ms / 1e3)

    # ---- decoder kernel alone (roofline)
    k_ms, fi, steps_exec = wl.decoder_leg(max(3, a.steps))
    roofline = wl.roofline(k_ms, fi, KERNEL_NAMES[a.schedule],
                           lookup_traffic(n, B, MAXITER, a.precision, a.schedule, a.lanes or auto_lanes(B), fi))

    # ---- end to end through the host-buffer C ABI
    e2e = None
    if not a.no_e2e:
        s2 = ClockSampler(local)          # clocks of THIS leg: it and the device leg are both power-capped, differently
        if rank == 0:
            s2.start()
            s2.wait_ready()
        el, h2d, d2h = wl.e2e_leg(a.steps, a.warmup, barrier, s2)
        e2e_clocks = s2.stop() if rank == 0 else None
        e2e = {"value": total_frames / max_over_ranks(el), "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "clocks": e2e_clocks,
               "api": "qr_reconcile_host: host y (f64) + tx symbols (i64) in; success, iterations, bit errors and "
                      "final LLRs out; pinned host memory"}

    # ---- the same through the compact wire format (an extension, own key: not the reference's array types)
    e2e_compact = None
    if not a.no_e2e:
        el, h2d, d2h = wl.e2e_compact_leg(a.steps, a.warmup, barrier)
        e2e_compact = {"value": total_frames / max_over_ranks(el), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h,
                       "api": "qr_reconcile_host_compact: host y (f32) + tx symbols (u8) in; success, iterations, bit errors "
                              "and packed hard decisions out; pinned host memory"}

    # ---- supplementary operating points (N = 1 only: they are reported, not scaled)
    points = None
    if world == 1 and not a.no_extras and n == N_CODE:
        points = extras(a, torch, qr, codes, dev, wl, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not a.no_cpu:
        cpu = cpu_baseline(a, torch, qr, wl)
    cnt = counters.cpu().tolist()
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "f64", "data": "synthetic",
            "config": workload_config(a, wl.n_sets),
            "info_gbit_per_s": value * K / 1e9,
            "avg_iterations": fi / B, "ber": cnt[0] / max(1, cnt[4] * K), "fer": cnt[1] / max(1, cnt[4]),
            "counters_all_reduced_inside_timed_region": True,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_compact": e2e_compact,
            # per step: k_front_end, k_eval_syndrome, k_demap32, k_init_batch, k_init_fused, k_fused, k_count_errors
            "gpu_launches": 7 * a.steps * world, "clocks": clocks}
    if points is not None:
        line["operating_points"] = points
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a, torch, qr, wl):
    """the CPU oracle decodes, one frame per process and call, the IDENTICAL frames taken from
    the GPU batch (input set 0); reported beside its agreement with the GPU's fp64 mode on those frames."""
    P = os.cpu_count() or 1
    fpw = a.cpu_frames or (2 if wl.n >= 30000 else 8)
    F = min(P * fpw, wl.B)
    y = wl.ys[0][:F].cpu().numpy(); x = wl.xs[0][:F].cpu().numpy()
    jobs = [[(y[f], x[f]) for f in range(w, F, P)] for w in range(P)]
    jobs = [j for j in jobs if j]
    arm = CpuArm("port", wl.n, a.snr, workers=len(jobs))
    try:
        r = arm.step(given=jobs)
    finally:
        arm.close()
    cpu_res = {}
    for w, lst in enumerate(r["results"]):
        for k, (ok, it, post) in enumerate(lst):
            cpu_res[w + k * len(jobs)] = (ok, it, post)
    # the same frames through the GPU's fp64 parity mode (exact bisection demapper, fused schedule where it applies)
    from qamreconciliation.pipeline import Reconciler
    rec64 = Reconciler(wl.dec, wl.nm, precision="fp64", demap="exact", lanes=min(512, (F + 31) // 32 * 32), schedule=3)
    out = rec64.run_device(wl.ys[0][:F], wl.xs[0][:F], MAXITER, k_info=wl.K)
    g_ok = out["success"].cpu().numpy(); g_it = out["iters"].cpu().numpy(); g_post = out["post"].cpu().numpy()
    same, rel_conv, dec_same = 0, 0.0, []
    for f in range(F):
        ok, it, post = cpu_res[f]
        eq = (ok == int(g_ok[f])) and (it == int(g_it[f]))
        same += eq
        if eq and ok:            # converged on both sides in the same iteration: compare the final LLRs
            rel_conv = max(rel_conv, float(np.max(np.abs(g_post[f] - post) / np.maximum(np.abs(post), 1e-300))))
        dec_same.append(float(np.mean((g_post[f] < 0) == (post < 0))))
    del rec64, out
    return {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
            "per_core_frames_per_s": r["fps"] / r["cores"], "info_gbit_per_s": r["fps"] * wl.K / 1e9,
            "edge_updates_per_s": r["fps"] * r["iters"] * wl.E,
            "sample": f"{r['cores']} processes x {len(jobs[0])} frames, the first {F} frames of the GPU batch (y, x copied to "
                      f"the host), one frame per process and call, whole chain (oracle/qr_oracle.c, gcc -O2); wall "
                      f"{r['wall']:.1f} s; per frame: front end+syndrome {r['chain']:.3f} s, demap {r['demap']:.2f} s, "
                      f"decode {r['dec']:.2f} s ({r['iters']:.0f} iterations)",
            "agreement": {"frames": F, "vs": "GPU fp64 mode (exact demapper) on the same frames",
                          "success_iters_equal": same / F,
                          "max_rel_llr_err_converged": rel_conv if same and any(cpu_res[f][0] for f in range(F)) else None,
                          "min_hard_decision_agreement": min(dec_same)}}


def extras(a, torch, qr, codes, dev, wl, barrier):
    """The same measurement at the operating points a reconciliation system runs at, the fp64 parity mode, and the
    decoder on BASELINE configs 3 and 4 -- in the driver-run line, so they are judged on driver-run numbers."""
    steps, warmup = min(a.steps, 5), 3
    pts = {}
    cfg = np.zeros(1 << BPS, dtype=np.uint8); cfg[1::2] = 1
    vid, cid = codes.regular_ldpc(a.n, DV, DC, seed=CODE_SEED)

    def measure(w, name, e2e=True, kname=None):
        ms, counters = w.device_leg(steps, warmup, barrier)
        cnt = counters.cpu().tolist()
        k_ms, fi, _ = w.decoder_leg(3)
        rec = {"value": w.B * steps / (ms / 1e3), "unit": "frames/s", "ms_per_step": ms / steps, "steps": steps,
               "warmup": warmup, "frames_per_step": w.B, "avg_iterations": fi / w.B,
               "fer": cnt[1] / max(1, cnt[4]), "ber": cnt[0] / max(1, cnt[4] * w.K),
               "roofline": w.roofline(k_ms, fi, kname or KERNEL_NAMES[w.schedule],
                                      lookup_traffic(w.n, w.B, w.maxiter, w.precision, w.schedule, w.lanes or auto_lanes(w.B), fi))}
        if e2e:
            el, h2d, d2h = w.e2e_leg(steps, warmup, barrier)
            rec["e2e"] = {"value": w.B * steps / el, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        pts[name] = rec

    for snr in (4.0, 5.0):
        w = Workload(torch, qr, codes, dev, 0, vid, cid, BPS, snr, cfg, a.frames, MAXITER, a.precision, a.demap, a.lanes,
                     a.schedule, dec=wl.dec)
        measure(w, f"config2_{snr:g}dB")
        del w
    w = Workload(torch, qr, codes, dev, 0, vid, cid, BPS, a.snr, cfg, 512, MAXITER, "fp64", "exact", 512, a.schedule,
                 dec=wl.dec)
    measure(w, "config2_fp64_parity_mode", e2e=False)
    del w
    torch.cuda.empty_cache()
    # BASELINE config 3: irregular R = 0.2, n = 131070 (3 * 43690), 8-PAM, 100 iterations max, below its waterfall
    v3, c3 = codes.irregular_ldpc(131070, 104856, [3, 8], [0.9, 0.1], seed=3)
    cfg3 = np.zeros(8, dtype=np.uint8); cfg3[1::2] = 1
    w = Workload(torch, qr, codes, dev, 0, v3, c3, 3, 3.0, cfg3, 1024, 100, "fp32", "fast", 1024, 3, n_sets=1)
    measure(w, "config3_irregular_n131070_8pam", e2e=False)
    del w
    torch.cuda.empty_cache()
    # BASELINE config 4: QKD scale, irregular R = 0.1, n = 2^20, 2-PAM, 60 iterations max
    v4, c4 = codes.irregular_ldpc(1 << 20, 943718, [3, 4, 10], [0.8, 0.15, 0.05], seed=4)
    w = Workload(torch, qr, codes, dev, 0, v4, c4, 1, -12.0, np.array([0, 1], dtype=np.uint8), 512, 60, "fp32", "fast",
                 512, 3, n_sets=1)
    measure(w, "config4_irregular_n1048576_2pam", e2e=False)
    del w
    torch.cuda.empty_cache()
    return pts


def workload_name(n, snr):
    return (f"(3,6)-regular LDPC n={n} R=1/2 (seed {CODE_SEED}), 4-PAM Alternating, Es/N0={snr} dB, "
            f"maxiter {MAXITER}, soft reverse reconciliation (BASELINE config 2)")


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
