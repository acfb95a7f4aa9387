#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref).

Run in THIS container only (needs oracle/_ref, i.e. /root/reference to have been built
by oracle/build_ref.py):

    python tests/golden/make_golden.py

The reference package is imported under its own name `qamreconciliation`, so this
script must not be run with the product package on sys.path; it loads the product's
code generator (codes.py, numpy only) by file path.  Inputs are seeded; outputs are
whatever the reference returned, stored verbatim (float64 bit patterns preserved).
The fixtures pin: the oracle restatement (CPU tests) and the CUDA path (GPU tests).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import qamreconciliation as ref  # noqa: E402  (the compiled reference)
from sims import reconciliation as refsims  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "qr_codes", os.path.join(ROOT, "qam-reconciliation_b200", "qamreconciliation", "codes.py"))
codes = importlib.util.module_from_spec(spec)
spec.loader.exec_module(codes)


def w(a, dtype):
    """Writable contiguous copy (the reference rejects read-only buffers)."""
    return np.ascontiguousarray(a, dtype=dtype).copy()


def arr(mv, dtype=None):
    a = np.asarray(mv)
    if a.dtype.kind == "S":          # cvarray format 'c'
        a = a.view(np.uint8)
    return np.array(a, dtype=dtype, copy=True)


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path)} bytes")


# ---------------------------------------------------------------- decoder KATs
def decoder_kats():
    vid, cid = codes.hamming_7_4()
    dec = ref.Decoder(w(vid, np.int64), w(cid, np.int64))
    out = dict(vid=vid, cid=cid)
    cases = {
        "kat1": ([1.2, -0.8, -1.3, 1.1, -0.4, 0.5, 1.9], [1, 1, 0]),
        "kat2": ([1.05, -1.075, -1.0, 1.1, -0.4, 0.4, -0.2], [1, 1, 0]),
        "noconv": ([0.3, -0.2, 0.1, -0.4, 0.2, -0.1, 0.05], [1, 0, 1]),
        "zeros": ([0.0, 0.0, -0.0, 0.0, 0.0, 0.0, 0.0], [0, 1, 0]),
    }
    for name, (llr, synd) in cases.items():
        for mi in (0, 1, 2, 20, 50):
            ok, it, post = dec.decode(w(llr, np.float64), w(synd, np.uint8), mi)
            out[f"{name}_llr"] = np.array(llr)
            out[f"{name}_synd"] = np.array(synd, dtype=np.uint8)
            out[f"{name}_m{mi}_ok"] = np.array(int(ok))
            out[f"{name}_m{mi}_it"] = np.array(int(it))
            out[f"{name}_m{mi}_post"] = arr(post)
    save("decoder_hamming.npz", **out)


def node_ops():
    """Single-node entry points on a small irregular graph (check degrees 2..7)."""
    rng = np.random.default_rng(11)
    cdeg = [2, 3, 4, 5, 6, 7, 3, 2]
    n = 12
    vid, cid = [], []
    for c, d in enumerate(cdeg):
        vs = rng.choice(n, size=d, replace=False)
        vid += list(vs); cid += [c] * d
    vid = np.array(vid, dtype=np.int64); cid = np.array(cid, dtype=np.int64)
    perm = rng.permutation(vid.size)          # unsorted edge list: association order matters
    vid, cid = vid[perm], cid[perm]
    dec = ref.Decoder(w(vid, np.int64), w(cid, np.int64))
    E, C, N = vid.size, len(cdeg), int(vid.max()) + 1
    v2c = rng.normal(0, 3, size=E); v2c[3] = 0.0; v2c[5] = 25.0; v2c[7] = -40.0
    synd = rng.integers(0, 2, size=C).astype(np.uint8)
    c2v = np.zeros(E)
    for c in range(C):
        dec.process_check_node(c, w(synd, np.uint8), c2v, w(v2c, np.float64))
    llr = rng.normal(0, 2, size=N)
    c2v_in = rng.normal(0, 2, size=E)
    v2c_out = np.zeros(E); post = np.zeros(N)
    for v in range(N):
        dec.process_var_node(v, w(llr, np.float64), w(c2v_in, np.float64), v2c_out, post)
    lap = rng.normal(0, 1, size=(16, N)); lap[0, :3] = 0.0
    syn = rng.integers(0, 2, size=(16, C)).astype(np.uint8)
    chk = np.array([int(dec.check_lappr(w(lap[i], np.float64), w(syn[i], np.uint8))) for i in range(16)])
    # make half of them consistent
    mat = ref.Matrix(w(vid, np.int64), w(cid, np.int64))
    for i in range(8):
        syn[i] = arr(mat.eval_syndrome(w((lap[i] < 0), np.uint8)))
    chk = np.array([int(dec.check_lappr(w(lap[i], np.float64), w(syn[i], np.uint8))) for i in range(16)])
    save("decoder_nodes.npz", vid=vid, cid=cid, v2c=v2c, synd=synd, c2v=c2v, llr=llr, c2v_in=c2v_in,
         v2c_out=v2c_out, post=post, lap=lap, syn=syn, chk=chk)


# ---------------------------------------------------------------- mapper
def mapper_cases():
    rng = np.random.default_rng(5)
    for bps in (1, 2, 3):
        pa = ref.PAMAlphabet(bps, 2)
        M = pa.order
        for cfg_name in ("base", "alt"):
            cfg = np.zeros(M, dtype=np.uint8)
            if cfg_name == "alt":
                cfg[1::2] = 1
            for snr in (2.0, 6.0):
                n0 = pa.variance * 10 ** (-snr / 10) / 2
                nm = ref.NoiseMapper(pa, n0, w(cfg, np.uint8))
                S = 160
                x = rng.integers(0, M, size=S).astype(np.int64)
                y = np.array(pa.constellation)[x] + np.sqrt(n0) * rng.normal(size=S)
                thr = np.array(pa.thresholds)
                special = np.concatenate([thr, thr[1:-1] - 1e-12, thr[1:-1] + 1e-12,
                                          [-1e3, 1e3, 0.0, -0.0, 40.0, -40.0]])
                y[:special.size] = special
                idx = arr(nm.hard_decide_index(w(y, np.float64)), np.int64)
                nh = arr(nm.map_noise(w(y, np.float64), w(idx, np.int64)))
                bits = arr(pa.demap_symbols_to_bits(w(idx, np.int64)), np.uint8)
                lap = arr(nm.demap_lappr_array(w(nh, np.float64), w(x, np.int64)))
                # every (n, j) pair on a grid, incl. n = 0 and 1
                ng = np.tile(np.array([0.0, 1e-12, 0.03, 0.25, 0.5, 0.77, 0.999, 1.0]), M)
                jg = np.repeat(np.arange(M), 8).astype(np.int64)
                lap_grid = arr(nm.demap_lappr_array(w(ng, np.float64), w(jg, np.int64)))
                yh_grid = np.array([[nm.g_inv_search(float(nv), i) for i in range(M)]
                                    for nv in ng[:8]])
                bare = arr(nm.bare_llr(w(x, np.int64)))
                direct = arr(refsims.y_to_lappr_grey_array(w(y, np.float64), pa, 2 * n0))
                save(f"mapper_bps{bps}_{cfg_name}_snr{int(snr)}.npz",
                     bps=bps, step=2.0, noise_var=n0, sign_config=cfg,
                     constellation=arr(pa.constellation), thresholds=arr(pa.thresholds),
                     probabilities=arr(pa.probabilities), variance=pa.variance,
                     s_to_b=arr(pa.s_to_b, np.uint8),
                     F_Y_thresholds=arr(nm.F_Y_thresholds), delta_F_Y=arr(nm.delta_F_Y),
                     fwrd=arr(nm.fwrd_transition_probability), back=arr(nm.back_transition_probability),
                     bare_llr_table=arr(nm.bare_llr_table), inf_erf_table=arr(nm.inf_erf_table),
                     x=x, y=y, idx=idx, n_hat=nh, bits=bits, lappr=lap, n_grid=ng, j_grid=jg,
                     lappr_grid=lap_grid, yhat_grid=yh_grid, bare=bare, direct=direct)


# ---------------------------------------------------------------- full chain
def chain_cases():
    """Whole frames through the reference, the way sims/reconciliation.pyx:127-153 chains them."""
    rng = np.random.default_rng(9)
    for (n, bps, snrs, frames, irregular) in ((96, 1, (1.0, 4.0), 4, False), (648, 2, (6.0, 9.0), 3, False),
                                              (648, 3, (12.0, 15.0), 2, False), (300, 2, (7.0, 10.0), 3, True)):
        if irregular:
            vid, cid = codes.irregular_ldpc(n, n // 2, [2, 3, 8], [0.5, 0.4, 0.1], seed=4)
        else:
            vid, cid = codes.regular_ldpc(n, 3, 6, seed=2)
        dec = ref.Decoder(w(vid, np.int64), w(cid, np.int64))
        mat = ref.Matrix(w(vid, np.int64), w(cid, np.int64))
        pa = ref.PAMAlphabet(bps, 2)
        M = pa.order
        cfg = np.zeros(M, dtype=np.uint8); cfg[1::2] = 1
        S = n // bps
        out = dict(vid=vid, cid=cid, bps=bps, sign_config=cfg, snrs=np.array(snrs), maxiter=50)
        for si, snr in enumerate(snrs):
            n0 = pa.variance * 10 ** (-snr / 10) / 2
            nm = ref.NoiseMapper(pa, n0, w(cfg, np.uint8))
            rec = {k: [] for k in ("x", "y", "x_hat", "n_hat", "word", "synd", "lappr", "ok", "it", "post",
                                   "hard_lappr", "hard_ok", "hard_it", "hard_post",
                                   "dir_word", "dir_synd", "dir_lappr", "dir_ok", "dir_it", "dir_post")}
            for _ in range(frames):
                x = rng.integers(0, M, size=S).astype(np.int64)
                y = np.array(pa.constellation)[x] + np.sqrt(n0) * rng.normal(size=S)
                xh = arr(nm.hard_decide_index(w(y, np.float64)), np.int64)
                nh = arr(nm.map_noise(w(y, np.float64), w(xh, np.int64)))
                word = arr(pa.demap_symbols_to_bits(w(xh, np.int64)), np.uint8)
                synd = arr(mat.eval_syndrome(w(word, np.uint8)), np.uint8)
                lap = arr(nm.demap_lappr_array(w(nh, np.float64), w(x, np.int64)))
                ok, it, post = dec.decode(w(lap, np.float64), w(synd, np.uint8), 50)
                rec["x"].append(x); rec["y"].append(y); rec["x_hat"].append(xh); rec["n_hat"].append(nh)
                rec["word"].append(word); rec["synd"].append(synd); rec["lappr"].append(lap)
                rec["ok"].append(int(ok)); rec["it"].append(int(it)); rec["post"].append(arr(post))
                # hard reverse (sims/reconciliation.pyx:300-308)
                hl = arr(nm.bare_llr(w(x, np.int64)))
                ok, it, post = dec.decode(w(hl, np.float64), w(synd, np.uint8), 50)
                rec["hard_lappr"].append(hl); rec["hard_ok"].append(int(ok)); rec["hard_it"].append(int(it))
                rec["hard_post"].append(arr(post))
                # soft direct (sims/reconciliation.pyx:214-227)
                dw = arr(pa.demap_symbols_to_bits(w(x, np.int64)), np.uint8)
                ds = arr(mat.eval_syndrome(w(dw, np.uint8)), np.uint8)
                dl = arr(refsims.y_to_lappr_grey_array(w(y, np.float64), pa, 2 * n0))
                ok, it, post = dec.decode(w(dl, np.float64), w(ds, np.uint8), 50)
                rec["dir_word"].append(dw); rec["dir_synd"].append(ds); rec["dir_lappr"].append(dl)
                rec["dir_ok"].append(int(ok)); rec["dir_it"].append(int(it)); rec["dir_post"].append(arr(post))
            for k, v in rec.items():
                out[f"s{si}_{k}"] = np.array(v)
            out[f"s{si}_noise_var"] = n0
        tag = "irr" if irregular else "reg"
        save(f"chain_{tag}_n{n}_bps{bps}.npz", **out)


if __name__ == "__main__":
    decoder_kats()
    node_ops()
    mapper_cases()
    chain_cases()
