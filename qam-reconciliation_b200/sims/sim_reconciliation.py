"""Command-line front end for the reconciliation simulations, flag-compatible with the reference's
sims/sim_reconciliation.py (:27-47 there) and writing the same CSV (EsN0dB, ber, fer, iters).

    python -m sims.sim_reconciliation EDGEFILE [--out out.csv] [--maxiter 50] [--ferr-count-min 100]
        [--alpha 1.0] [--simloops 5000] [--snr 0 5] [--nsnr 11] [--bps 2] [--hard] [--direct]
        [--configuration-base]

The reference's own script runs unchanged on this package as well; this one
exists so the repository has a front end of its own.  Under torchrun every SNR point's frames are
split across the GPUs and rank 0 writes the file.
"""
import argparse
import os


def main(argv=None):
    import numpy as np
    import pandas as pd
    import torch

    from qamreconciliation import Decoder, Matrix, PAMAlphabet
    from sims.reconciliation import (simulate_direct_snr_dB, simulate_hard_reverse_snr_dB,
                                     simulate_softening_snr_dB)

    ap = argparse.ArgumentParser(prog="sim_reconciliation", description="BER / FER / iterations of LDPC-based "
                                 "reconciliation over PAM, soft reverse (default), hard reverse or soft direct")
    ap.add_argument("edgefile", help="CSV with eid,cid,vid columns; first data row holds the counts E,C,N")
    ap.add_argument("--out", default="out.csv")
    ap.add_argument("--maxiter", default=50, type=int)
    ap.add_argument("--ferr-count-min", default=100, type=int)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--simloops", default=5000, type=int)
    ap.add_argument("--snr", type=float, nargs=2, default=[0, 5])
    ap.add_argument("--nsnr", type=int, default=11)
    ap.add_argument("--bps", type=int, default=2)
    ap.add_argument("--hard", action="store_true")
    ap.add_argument("--direct", action="store_true")
    ap.add_argument("--configuration-base", action="store_true")
    a = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    edges = pd.read_csv(a.edgefile)
    vid = edges.vid[1:].to_numpy(); cid = edges.cid[1:].to_numpy()
    dec = Decoder(vid, cid); mat = Matrix(vid, cid); pa = PAMAlphabet(a.bps, 2)
    rows = []
    for snr in np.linspace(a.snr[0], a.snr[1], a.nsnr):
        if a.direct:
            rows.append(simulate_direct_snr_dB(snr, dec, mat, pa, a.maxiter, a.simloops, a.ferr_count_min))
        elif a.hard:
            rows.append(simulate_hard_reverse_snr_dB(snr, dec, mat, pa, a.maxiter, a.simloops, a.ferr_count_min))
        else:
            cfg = np.zeros(pa.order, dtype=np.uint8)
            if not a.configuration_base:
                cfg[1::2] = 1
            rows.append(simulate_softening_snr_dB(snr, dec, mat, pa, cfg, a.maxiter, a.simloops,
                                                  a.ferr_count_min, a.alpha))
    if int(os.environ.get("RANK", "0")) == 0:
        pd.DataFrame(rows, columns=["EsN0dB", "ber", "fer", "iters"]).to_csv(a.out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
