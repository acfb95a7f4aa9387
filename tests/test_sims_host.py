"""Host-side logic of the batched Monte-Carlo drivers: the sequential statistics of the reference's
loop (sims/reconciliation.pyx:127-168) reproduced on batched per-frame results, and the multi-rank
exchange (world_size 2 over gloo on the CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qam-reconciliation_b200")


def reference_loop(errors, success, iters, simulation_loops, ferr_count_min, K):
    """The reference's loop, frame by frame (sims/reconciliation.pyx:114-168), on precomputed outcomes."""
    err_count = frame_error_count = decoding_iterations = successful_decoding = 0
    wordcount = 0
    for wordcount in range(simulation_loops):
        if success[wordcount]:
            decoding_iterations += iters[wordcount]
            successful_decoding += 1
        new_errors = errors[wordcount]
        if new_errors:
            frame_error_count += 1
            err_count += new_errors
        if frame_error_count >= ferr_count_min and wordcount > simulation_loops / 20:
            break
    wordcount += 1
    return (err_count / (wordcount * K), frame_error_count / wordcount,
            0 if successful_decoding == 0 else decoding_iterations / successful_decoding, wordcount)


def batched(errors, success, iters, simulation_loops, ferr_count_min, K, batch):
    from sims.reconciliation import sequential_statistics
    st = dict(bit_errors=0, frame_errors=0, successes=0, iterations=0, frames=0)
    pos = 0
    done = False
    while not done:
        sl = slice(pos, pos + batch)
        done = sequential_statistics(errors[sl], success[sl], iters[sl], simulation_loops, ferr_count_min, st)
        pos += batch
    return (st["bit_errors"] / (st["frames"] * K), st["frame_errors"] / st["frames"],
            0 if st["successes"] == 0 else st["iterations"] / st["successes"], st["frames"])


@pytest.mark.parametrize("seed", range(6))
def test_sequential_statistics_equal_the_reference_loop(seed):
    rng = np.random.default_rng(seed)
    loops = int(rng.integers(30, 400))
    p_err = rng.choice([0.0, 0.02, 0.3, 0.9])
    n = loops + 700
    errors = np.where(rng.random(n) < p_err, rng.integers(1, 40, n), 0)
    success = (rng.random(n) < 0.7).astype(np.int64)
    iters = rng.integers(1, 51, n)
    for fmin in (1, 5, 100):
        want = reference_loop(errors, success, iters, loops, fmin, 324)
        for batch in (1, 7, 64, 1000):
            assert batched(errors, success, iters, loops, fmin, 324, batch) == want


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, q):
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sims.reconciliation import gather_frame_stats, sequential_statistics
    rng = np.random.default_rng(100 + rank)
    st = dict(bit_errors=0, frame_errors=0, successes=0, iterations=0, frames=0)
    done, rounds = False, 0
    local_log = []
    while not done:
        local = np.stack([np.where(rng.random(16) < 0.2, rng.integers(1, 9, 16), 0), rng.integers(0, 2, 16),
                          rng.integers(1, 51, 16)], axis=1).astype(np.int32)
        local_log.append(local)
        allst = gather_frame_stats(torch.from_numpy(local)).numpy()
        assert allst.shape == (16 * world, 3)
        assert np.array_equal(allst[16 * rank:16 * (rank + 1)], local)      # rank-major order
        done = sequential_statistics(allst[:, 0], allst[:, 1], allst[:, 2], 200, 12, st)
        rounds += 1
    q.put((rank, st, rounds, np.concatenate(local_log)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_gives_every_rank_the_same_statistics():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, st0, r0, log0), (_, st1, r1, log1) = res
    assert st0 == st1 and r0 == r1
    # and they equal the one-process loop over the interleaved (rank-major per batch) frame order
    order = np.concatenate([np.concatenate([log0[16 * b:16 * (b + 1)], log1[16 * b:16 * (b + 1)]]) for b in range(r0)])
    sys.path.insert(0, PKG)
    want = reference_loop(order[:, 0], order[:, 1], order[:, 2], 200, 12, 1)
    assert st0["frames"] == want[3] and st0["frame_errors"] / st0["frames"] == want[1]


def test_parfor_stand_in():
    sys.path.insert(0, PKG)
    from parfor import parfor

    @parfor([1.0, 2.0, 3.0])
    def results(x):
        return (x, x * x)
    assert results == [(1.0, 1.0), (2.0, 4.0), (3.0, 9.0)]
