import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qam-reconciliation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qamreconciliation as qr
from qamreconciliation import codes
from oracle import port as orc
from test_gpu_parity import bpsk_frames
vid, cid = codes.regular_ldpc(6480, 3, 6, seed=1)
frames = 96
word, llr, synd = bpsk_frames(orc, vid, cid, frames, [0.80, 0.84], seed=4)
ook, oit, opost = orc.Decoder(vid, cid).decode_frames(llr, synd, 50)
dec = qr.Decoder(vid, cid)
for lanes in (None, 96, 128, 672):
    for schedule in (0, 1):
        ok, it, post = dec.decode_batch(torch.tensor(llr, dtype=torch.float32), synd, 50, precision="fp32", schedule=schedule, lanes=lanes)
        ok, it, post = ok.cpu().numpy(), it.cpu().numpy(), post.cpu().numpy()
        both = (ok == 1) & (ook == 1)
        diff = ((post < 0) != (opost < 0)).sum(axis=1)
        werr = ((post < 0).astype(np.uint8) != word).sum(axis=1)
        print("lanes", lanes, "sched", schedule, "ok agree", (ok == ook).mean(), "frames with decision diff:", np.flatnonzero(diff[both] > 0)[:10], diff[both][diff[both] > 0][:10],
              "word errs", werr[both][werr[both] > 0][:10], "it", it[:8], oit[:8])
        bad = np.flatnonzero((diff > 0) & both)
        for f in bad[:3]:
            cols = np.flatnonzero((post[f] < 0) != (opost[f] < 0))
            print("  frame", f, "it", it[f], oit[f], "cols", cols[:10], post[f, cols[:5]], opost[f, cols[:5]])
