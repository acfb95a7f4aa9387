"""Batched Monte-Carlo drivers with the reference's entry points (sims/reconciliation.pyx).

The reference binds to its extension classes through Cython `cimport` (reconciliation.pyx:17-18,
calling the cdef `Decoder._decode`), which a non-Cython class cannot satisfy, so the three drivers
are re-provided here with the same names, arguments and return value
    (snr_dB, ber, fer, average iterations over successful frames)
(reconciliation.pyx:93-96, :165-168, :173-176, :253-256).

Frames are simulated in batches on the GPU: symbols and noise are drawn on the device, the whole
chain hard decision -> softening metric -> bits -> syndrome -> LLR -> decoding -> error count runs
batched, and only three small integers per frame come back.  The reference's statistics are
sequential; they are reproduced exactly on the batched results:
  * a frame counts as a frame error iff it has a bit error in its first K = N - C bits, whatever the
    decoder's success flag says (reconciliation.pyx:153-156);
  * iterations are averaged over successful frames only, 0 if there are none (:149-151, :168);
  * the loop stops after frame w (0-based) as soon as frame_errors >= ferr_count_min and
    w > simulation_loops / 20 (:159-161); frames simulated beyond w in the same batch are discarded.
With torch.distributed initialised (one process per GPU) every batch is split across the ranks and
the per-frame counters are all-gathered (12 bytes per frame), so every rank takes the same stop
decision and returns the same tuple.

Environment knobs (the reference's CLI stays byte-identical): QAMRECON_PRECISION=fp32|fp64
(default fp32 here), QAMRECON_DEMAP=fast|exact, QAMRECON_SIM_BATCH (frames per rank per batch).
"""
import math
import os

import numpy as np
import torch

from qamreconciliation import NoiseMapper
from qamreconciliation.pipeline import HARD_REVERSE, SOFT_DIRECT, SOFT_REVERSE, Reconciler


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _precision():
    p = os.environ.get("QAMRECON_PRECISION", "fp32").lower()
    return "fp64" if p in ("fp64", "f64", "64", "double") else "fp32"


def gather_frame_stats(stats):
    """Per-frame counters [frames_per_rank, 3] (bit errors, success, iterations) of every rank, rank-major:
    the only exchange of the path (12 bytes per frame; NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    dist, rank, world = _dist()
    if dist is None or world == 1:
        return stats
    gathered = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)
    return torch.cat(gathered, dim=0)


def sequential_statistics(errors, success, iters, simulation_loops, ferr_count_min, state):
    """Fold a batch of per-frame results (global frame order) into the running counters exactly as
    the reference's loop would, one frame after the other.  Returns True when the loop breaks."""
    errors = np.asarray(errors, dtype=np.int64); success = np.asarray(success, dtype=np.int64)
    iters = np.asarray(iters, dtype=np.int64)
    n = errors.size
    w0 = state["frames"]
    room = simulation_loops - w0
    if n > room:
        errors, success, iters, n = errors[:room], success[:room], iters[:room], room
    ferr_cum = state["frame_errors"] + np.cumsum(errors > 0)
    idx = w0 + np.arange(n)
    stop = np.flatnonzero((ferr_cum >= ferr_count_min) & (idx > simulation_loops / 20))
    last = int(stop[0]) if stop.size else n - 1
    take = slice(0, last + 1)
    state["bit_errors"] += int(errors[take].sum())
    state["frame_errors"] += int((errors[take] > 0).sum())
    state["successes"] += int(success[take].sum())
    state["iterations"] += int((iters[take] * success[take]).sum())
    state["frames"] += last + 1
    return bool(stop.size) or state["frames"] >= simulation_loops


def _simulate(snr_dB, dec, mat, pa, nmconfig, decoder_iterations, simulation_loops, ferr_count_min, alpha, mode):
    dist, rank, world = _dist()
    Es = pa.variance
    N0 = Es * (10 ** (-snr_dB / 10)) / 2            # reconciliation.pyx:109-110, :195-197
    nm = NoiseMapper(pa, N0, nmconfig) if nmconfig is not None else NoiseMapper(pa, N0)
    N = mat.vnum
    K = N - mat.cnum
    if N % pa.bit_per_symbol:
        raise ValueError(f"codeword length {N} is not a multiple of bits per symbol {pa.bit_per_symbol}")
    S = N // pa.bit_per_symbol
    rec = Reconciler(dec, nm, mode=mode, precision=_precision(), demap=os.environ.get("QAMRECON_DEMAP", "fast"),
                     alpha=alpha)
    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev)
    seed = int(np.random.randint(0, 2 ** 31 - 1))   # follows np.random.seed(), like the reference's global RNG
    if dist is not None:
        t = torch.tensor([seed], device=dev)
        dist.broadcast(t, 0)
        seed = int(t.item())
    gen.manual_seed(seed * 131 + rank)
    const = torch.as_tensor(np.asarray(pa.constellation), dtype=torch.float64, device=dev)
    probs = torch.as_tensor(np.asarray(pa.probabilities), dtype=torch.float64, device=dev)
    per_rank = int(os.environ.get("QAMRECON_SIM_BATCH", "0")) or max(1, min(2048, -(-simulation_loops // world)))
    state = dict(bit_errors=0, frame_errors=0, successes=0, iterations=0, frames=0)
    sigma = math.sqrt(N0)
    done = simulation_loops <= 0
    while not done:
        x = torch.multinomial(probs, per_rank * S, replacement=True, generator=gen).reshape(per_rank, S)
        y = const[x] + sigma * torch.randn((per_rank, S), dtype=torch.float64, device=dev, generator=gen)
        out = rec.run_device(y, x, int(decoder_iterations), k_info=K, want_post=False)
        stats = torch.stack([out["bit_errors"], out["success"].to(torch.int32), out["iters"]], dim=1).contiguous()
        s = gather_frame_stats(stats).cpu().numpy()
        done = sequential_statistics(s[:, 0], s[:, 1], s[:, 2], simulation_loops, ferr_count_min, state)
    frames = max(state["frames"], 1)
    return (snr_dB,
            state["bit_errors"] / (frames * K),
            state["frame_errors"] / frames,
            0 if state["successes"] == 0 else state["iterations"] / state["successes"])


def simulate_softening_snr_dB(snr_dB, dec, mat, pa, nmconfig, decoder_iterations, simulation_loops,
                              ferr_count_min, alpha=1.0):
    """Soft reverse reconciliation (reconciliation.pyx:93-168)."""
    return _simulate(snr_dB, dec, mat, pa, nmconfig, decoder_iterations, simulation_loops, ferr_count_min, alpha,
                     SOFT_REVERSE)


def simulate_direct_snr_dB(snr_dB, dec, mat, pa, decoder_iterations, simulation_loops, ferr_count_min):
    """Soft direct reconciliation (reconciliation.pyx:173-249)."""
    return _simulate(snr_dB, dec, mat, pa, None, decoder_iterations, simulation_loops, ferr_count_min, 1.0,
                     SOFT_DIRECT)


def simulate_hard_reverse_snr_dB(snr_dB, dec, mat, pa, decoder_iterations, simulation_loops, ferr_count_min):
    """Hard reverse reconciliation (reconciliation.pyx:253-329)."""
    return _simulate(snr_dB, dec, mat, pa, None, decoder_iterations, simulation_loops, ferr_count_min, 1.0,
                     HARD_REVERSE)


def y_to_lappr_grey_array(y, pa, twoVariance):
    """reconciliation.pyx:75-90: direct-reconciliation LLRs of the samples y (numpy in, numpy out)."""
    nm = NoiseMapper(pa, twoVariance / 2)
    return nm.direct_llr_batch(np.asarray(y, dtype=np.float64).reshape(-1), two_variance=twoVariance).cpu().numpy()
