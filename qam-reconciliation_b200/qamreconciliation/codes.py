"""Seeded synthetic LDPC edge lists and the reference's edge-list CSV format.

The reference ships one 12-edge fixture (test/hamming_7-4.csv) and reads codes as a
CSV with header `eid,cid,vid`, a first data row holding the counts (E, C, N) and one
edge per following row (sims/sim_reconciliation.py:50-51,60 of the reference).  The
BASELINE configs need codes the reference does not ship, so they are generated here,
deterministically from a seed.

Host-side numpy only: this is code ingestion, not the hot path.
"""
import numpy as np


def _repair_duplicates(vsock, csock, rng, max_rounds=200):
    """Swap variable sockets until no (check, variable) pair occurs twice."""
    E = vsock.size
    for _ in range(max_rounds):
        key = csock.astype(np.int64) * (int(vsock.max()) + 1) + vsock
        order = np.argsort(key, kind="stable")
        dup = np.zeros(E, dtype=bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        bad = np.flatnonzero(dup)
        if bad.size == 0:
            return vsock
        partners = rng.integers(0, E, size=bad.size)
        for b, p in zip(bad, partners):
            vsock[b], vsock[p] = vsock[p], vsock[b]
    raise RuntimeError("could not remove duplicate edges")


def _finish(vsock, csock):
    order = np.lexsort((vsock, csock))          # sorted by (cid, vid)
    return vsock[order].astype(np.int64), csock[order].astype(np.int64)


def regular_ldpc(n, dv=3, dc=6, seed=1):
    """(dv,dc)-regular code: dv layers, each a seeded permutation of the variables dealt
    dc/dv per check.  Returns (vid, cid), sorted by (cid, vid)."""
    if (n * dv) % dc or dc % dv:
        raise ValueError("need dc | n*dv and dv | dc")
    c = n * dv // dc
    per = dc // dv
    rng = np.random.default_rng(seed)
    vs, cs = [], []
    for _ in range(dv):
        vs.append(rng.permutation(n))
        cs.append(np.repeat(np.arange(c), per))
    vsock = np.concatenate(vs); csock = np.concatenate(cs)
    vsock = _repair_duplicates(vsock, csock, rng)
    return _finish(vsock, csock)


def irregular_ldpc(n, c, var_degrees, var_fractions, seed=1):
    """Irregular code by the configuration model: variable degrees drawn to match the given
    node-perspective fractions, check degrees as even as possible (all >= 2).  Returns (vid, cid)."""
    rng = np.random.default_rng(seed)
    var_degrees = np.asarray(var_degrees, dtype=np.int64)
    frac = np.asarray(var_fractions, dtype=np.float64)
    counts = np.floor(frac / frac.sum() * n).astype(np.int64)
    counts[0] += n - counts.sum()
    dv = np.repeat(var_degrees, counts)
    rng.shuffle(dv)
    E = int(dv.sum())
    base, extra = divmod(E, c)
    if base < 2:
        raise ValueError("average check degree below 2")
    dcv = np.full(c, base, dtype=np.int64)
    dcv[rng.permutation(c)[:extra]] += 1
    vsock = np.repeat(np.arange(n), dv)
    rng.shuffle(vsock)
    csock = np.repeat(np.arange(c), dcv)
    vsock = _repair_duplicates(vsock, csock, rng)
    return _finish(vsock, csock)


def hamming_7_4():
    """The Hamming(7,4) graph of the reference's fixture test/hamming_7-4.csv (12 edges)."""
    cid = np.array([0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2], dtype=np.int64)
    vid = np.array([0, 3, 4, 6, 1, 3, 5, 6, 2, 4, 5, 6], dtype=np.int64)
    return vid, cid


def write_edge_csv(path, vid, cid):
    """Reference format: header, one row of counts (E, C, N), then `eid, cid, vid` rows."""
    vid = np.asarray(vid); cid = np.asarray(cid)
    with open(path, "w") as fh:
        fh.write("eid,cid,vid\n")
        fh.write(f"{vid.size},{int(cid.max()) + 1},{int(vid.max()) + 1}\n")
        for e, (c, v) in enumerate(zip(cid, vid)):
            fh.write(f"{e},{int(c)},{int(v)}\n")


def read_edge_csv(path):
    """Inverse of write_edge_csv; returns (vid, cid) without the counts row."""
    data = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    return data[1:, 2].copy(), data[1:, 1].copy()
