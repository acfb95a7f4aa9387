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


# ---------------------------------------------------------------------------------------------
# Code ingestion either side of the hot path

def load_edge_csv(path):
    """Edge-list CSV of the reference (test/hamming_7-4.csv; sims/sim_reconciliation.py:50-51,60): header
    `eid,cid,vid` in any column order, first data row = counts (E, C, N), one edge per following row, blanks
    after commas allowed.  Returns (vid, cid) in file order and checks them against the counts row."""
    with open(path) as fh:
        header = [h.strip() for h in fh.readline().split(",")]
    if sorted(header) != ["cid", "eid", "vid"]:
        raise ValueError(f"{path}: header must name the columns eid, cid, vid (got {header})")
    data = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    col = {name: i for i, name in enumerate(header)}
    counts, body = data[0], data[1:]
    vid, cid = body[:, col["vid"]].copy(), body[:, col["cid"]].copy()
    E, Cn, N = int(counts[col["eid"]]), int(counts[col["cid"]]), int(counts[col["vid"]])
    if vid.size != E:
        raise ValueError(f"{path}: counts row announces {E} edges, file holds {vid.size}")
    if vid.size and (int(cid.max()) + 1 != Cn or int(vid.max()) + 1 != N):
        raise ValueError(f"{path}: counts row announces {Cn} checks / {N} variables, edges span "
                         f"{int(cid.max()) + 1} / {int(vid.max()) + 1}")
    return vid, cid


def validate_edges(vid, cid):
    """What the reference leaves unchecked: returns a dict with the number of duplicate
    edges, the checks of degree < 2 (the reference reads out of bounds for them, decoder.pyx:131-135), unused
    variable ids (gaps), negative ids, and the degree spectra.  Raises nothing: the caller decides."""
    vid = np.asarray(vid, dtype=np.int64).ravel(); cid = np.asarray(cid, dtype=np.int64).ravel()
    if vid.size != cid.size:
        raise ValueError("Sizes don't match")
    rep = {"edges": int(vid.size), "negative_ids": int((vid < 0).sum() + (cid < 0).sum())}
    if vid.size == 0 or rep["negative_ids"]:
        rep.update(duplicate_edges=0, weak_checks=[], unused_variables=[], check_degrees={}, variable_degrees={})
        return rep
    N, Cn = int(vid.max()) + 1, int(cid.max()) + 1
    key = np.sort(cid * N + vid)
    rep["duplicate_edges"] = int((key[1:] == key[:-1]).sum())
    cdeg, vdeg = np.bincount(cid, minlength=Cn), np.bincount(vid, minlength=N)
    rep["weak_checks"] = np.flatnonzero(cdeg < 2).tolist()
    rep["unused_variables"] = np.flatnonzero(vdeg == 0).tolist()
    rep["check_degrees"] = {int(d): int(c) for d, c in zip(*np.unique(cdeg, return_counts=True))}
    rep["variable_degrees"] = {int(d): int(c) for d, c in zip(*np.unique(vdeg, return_counts=True))}
    rep["variables"], rep["checks"] = N, Cn
    return rep


def read_alist(path):
    """MacKay alist parity-check format -> (vid, cid), edges ordered by (check, position in the check's row).
    Layout: `N M` / `max_col_w max_row_w` / N column weights / M row weights / N column lists / M row lists,
    1-based, zero-padded.  The row lists are authoritative; the column lists are cross-checked."""
    with open(path) as fh:
        tok = fh.read().split()
    it = iter(int(t) for t in tok)
    try:
        N, M = next(it), next(it)
        next(it), next(it)
        colw = [next(it) for _ in range(N)]
        roww = [next(it) for _ in range(M)]
        cols = []
        total = list(it)
    except StopIteration:
        raise ValueError(f"{path}: truncated alist header")
    # the lists may be zero-padded to the maximum weight or not; consume by trying padded first
    def take(lists_w, maxw, data, pos, padded):
        out = []
        for wgt in lists_w:
            n = maxw if padded else wgt
            seg = data[pos:pos + n]
            if len(seg) < n:
                return None, pos
            out.append([x for x in seg if x != 0])
            if len(out[-1]) != wgt:
                return None, pos
            pos += n
        return out, pos
    maxc, maxr = max(colw) if colw else 0, max(roww) if roww else 0
    for padded in (True, False):
        cols, pos = take(colw, maxc, total, 0, padded)
        if cols is None:
            continue
        rows, pos2 = take(roww, maxr, total, pos, padded)
        if rows is not None and pos2 == len(total):
            break
    else:
        raise ValueError(f"{path}: alist body does not match the announced weights")
    vid = np.array([v - 1 for r in rows for v in r], dtype=np.int64)
    cid = np.array([c for c, r in enumerate(rows) for _ in r], dtype=np.int64)
    from_cols = sorted((c - 1, v) for v, cl in enumerate(cols) for c in cl)
    if from_cols != sorted(zip(cid.tolist(), vid.tolist())):
        raise ValueError(f"{path}: column lists and row lists describe different matrices")
    return vid, cid


def write_alist(path, vid, cid):
    vid = np.asarray(vid, dtype=np.int64); cid = np.asarray(cid, dtype=np.int64)
    N, M = int(vid.max()) + 1, int(cid.max()) + 1
    cols = [[] for _ in range(N)]; rows = [[] for _ in range(M)]
    for v, c in zip(vid.tolist(), cid.tolist()):
        cols[v].append(c + 1); rows[c].append(v + 1)
    maxc, maxr = max(map(len, cols)), max(map(len, rows))
    with open(path, "w") as fh:
        fh.write(f"{N} {M}\n{maxc} {maxr}\n")
        fh.write(" ".join(str(len(c)) for c in cols) + "\n")
        fh.write(" ".join(str(len(r)) for r in rows) + "\n")
        for c in cols:
            fh.write(" ".join(map(str, c + [0] * (maxc - len(c)))) + "\n")
        for r in rows:
            fh.write(" ".join(map(str, r + [0] * (maxr - len(r)))) + "\n")


def edges_digest(vid, cid):
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(vid, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(cid, dtype=np.int64).tobytes())
    return h.hexdigest()


def save_tables(path, vid, cid, tables):
    """Cache of the check-grouped (CSR) / variable-grouped (CSC) permutations the library builds
    (qr_graph_export: chk_order, slot_edge, slot_var, var_ptr, var_slot), keyed by a digest of the edge list."""
    np.savez_compressed(path, digest=edges_digest(vid, cid), **tables)


def load_tables(path, vid, cid):
    """The cached tables, or None if the file belongs to another edge list."""
    z = np.load(path)
    if str(z["digest"]) != edges_digest(vid, cid):
        return None
    return {k: z[k] for k in z.files if k != "digest"}


def build_tables(vid, cid):
    """The library's O(E) graph build (csrc/qr_graph_build.h, replacing the reference's O(nodes * E) scan,
    decoder.pyx:60-89) on the host only: returns the CSR / CSC permutations as numpy arrays.  Raises
    ValueError for what the library rejects (negative ids, a check of degree < 2, degree > 64)."""
    import ctypes as C
    from . import _abi
    vid = np.ascontiguousarray(vid, dtype=np.int64).ravel(); cid = np.ascontiguousarray(cid, dtype=np.int64).ravel()
    if vid.size != cid.size:
        raise ValueError("Sizes don't match")
    h = C.c_void_p()
    _abi.check(_abi.lib().qr_graph_create(vid.ctypes.data, cid.ctypes.data, vid.size, -1, C.byref(h)))
    try:
        n, c, e = C.c_int64(), C.c_int64(), C.c_int64()
        _abi.check(_abi.lib().qr_graph_info(h, C.byref(n), C.byref(c), C.byref(e), None, None))
        t = dict(chk_order=np.zeros(c.value, np.int32), slot_edge=np.zeros(e.value, np.int32),
                 slot_var=np.zeros(e.value, np.int32), var_ptr=np.zeros(n.value + 1, np.int32),
                 var_slot=np.zeros(e.value, np.int32))
        _abi.check(_abi.lib().qr_graph_export(h, *(t[k].ctypes.data for k in
                                                   ("chk_order", "slot_edge", "slot_var", "var_ptr", "var_slot"))))
        return t
    finally:
        _abi.lib().qr_graph_destroy(h)
