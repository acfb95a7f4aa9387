"""Code ingestion: the reference's edge-list CSV, validation of what the
reference leaves unchecked, alist import/export, the O(E) table build and its on-disk cache.  CPU only."""
import os

import numpy as np
import pytest

from qamreconciliation import codes


def test_reference_csv_fixture_layout(tmp_path):
    # the reference's own fixture layout (test/hamming_7-4.csv: blanks after commas, counts row first)
    p = tmp_path / "h.csv"
    vid, cid = codes.hamming_7_4()
    with open(p, "w") as fh:
        fh.write("eid,cid,vid\n12,    3,     7\n")
        for e, (c, v) in enumerate(zip(cid, vid)):
            fh.write(f"{e},     {c},     {v}\n")
    v2, c2 = codes.load_edge_csv(p)
    assert np.array_equal(v2, vid) and np.array_equal(c2, cid)
    v3, c3 = codes.read_edge_csv(p)
    assert np.array_equal(v3, vid) and np.array_equal(c3, cid)
    # column order is taken from the header
    q = tmp_path / "h2.csv"
    with open(q, "w") as fh:
        fh.write("vid,eid,cid\n7,12,3\n")
        for e, (c, v) in enumerate(zip(cid, vid)):
            fh.write(f"{v},{e},{c}\n")
    v4, c4 = codes.load_edge_csv(q)
    assert np.array_equal(v4, vid) and np.array_equal(c4, cid)
    bad = tmp_path / "bad.csv"
    bad.write_text("eid,cid,vid\n11,3,7\n" + "".join(f"{e},{c},{v}\n" for e, (c, v) in enumerate(zip(cid, vid))))
    with pytest.raises(ValueError):
        codes.load_edge_csv(bad)


def test_csv_round_trip_large(tmp_path):
    vid, cid = codes.regular_ldpc(648, 3, 6, seed=3)
    p = tmp_path / "c.csv"
    codes.write_edge_csv(p, vid, cid)
    v, c = codes.load_edge_csv(p)
    assert np.array_equal(v, vid) and np.array_equal(c, cid)


def test_validate_edges_reports_what_the_reference_ignores():
    vid = np.array([0, 1, 1, 2, 4, 4]); cid = np.array([0, 0, 0, 1, 2, 2])
    rep = codes.validate_edges(vid, cid)
    assert rep["duplicate_edges"] == 2            # (0,1) and (2,4) twice
    assert rep["weak_checks"] == [1]              # degree 1: the reference reads out of bounds there
    assert rep["unused_variables"] == [3]         # id gap
    assert rep["check_degrees"] == {1: 1, 2: 1, 3: 1}
    assert codes.validate_edges([0, -1], [0, 0])["negative_ids"] == 1
    ok = codes.validate_edges(*codes.regular_ldpc(96, 3, 6, seed=1))
    assert ok["duplicate_edges"] == 0 and not ok["weak_checks"] and not ok["unused_variables"]
    assert ok["check_degrees"] == {6: 48} and ok["variable_degrees"] == {3: 96}


@pytest.mark.parametrize("gen", ["hamming", "regular", "irregular"])
def test_alist_round_trip(tmp_path, gen):
    vid, cid = {"hamming": codes.hamming_7_4, "regular": lambda: codes.regular_ldpc(96, 3, 6, seed=2),
                "irregular": lambda: codes.irregular_ldpc(120, 60, [2, 3, 7], [0.5, 0.4, 0.1], seed=5)}[gen]()
    p = tmp_path / "c.alist"
    codes.write_alist(p, vid, cid)
    v, c = codes.read_alist(p)
    assert sorted(zip(c.tolist(), v.tolist())) == sorted(zip(cid.tolist(), vid.tolist()))
    assert np.all(np.diff(c) >= 0)                # edges ordered by check
    # an inconsistent file is rejected
    txt = p.read_text().split("\n")
    txt[-2] = txt[-2].replace(txt[-2].split()[0], "1", 1) if txt[-2].split()[0] != "1" else txt[-2].replace("1", "2", 1)
    (tmp_path / "bad.alist").write_text("\n".join(txt))
    with pytest.raises(ValueError):
        codes.read_alist(tmp_path / "bad.alist")


def test_unpadded_alist(tmp_path):
    # Hamming(7,4) written WITHOUT zero padding
    p = tmp_path / "h.alist"
    p.write_text("7 3\n3 4\n1 1 1 2 2 2 3\n4 4 4\n1\n2\n3\n1 2\n1 3\n2 3\n1 2 3\n1 4 5 7\n2 4 6 7\n3 5 6 7\n")
    v, c = codes.read_alist(p)
    assert v.size == 12 and c.max() == 2 and v.max() == 6
    hv, hc = codes.hamming_7_4()
    assert sorted(zip(c.tolist(), v.tolist())) == sorted(zip(hc.tolist(), hv.tolist()))


def test_table_build_and_cache(tmp_path):
    rng = np.random.default_rng(0)
    vid, cid = codes.irregular_ldpc(300, 150, [2, 3, 8], [0.5, 0.4, 0.1], seed=4)
    perm = rng.permutation(vid.size)
    vid, cid = vid[perm], cid[perm]
    t = codes.build_tables(vid, cid)
    assert sorted(t["slot_edge"].tolist()) == list(range(vid.size))
    assert np.array_equal(t["slot_var"], vid[t["slot_edge"]])
    # per variable: ascending original edge id (decoder.pyx:73-76)
    for v in range(300):
        ed = t["slot_edge"][t["var_slot"][t["var_ptr"][v]:t["var_ptr"][v + 1]]]
        assert np.all(vid[ed] == v) and np.all(np.diff(ed) > 0)
    p = tmp_path / "tables.npz"
    codes.save_tables(p, vid, cid, t)
    back = codes.load_tables(p, vid, cid)
    assert all(np.array_equal(back[k], t[k]) for k in t)
    assert codes.load_tables(p, vid[::-1].copy(), cid[::-1].copy()) is None     # another edge list
    with pytest.raises(ValueError):
        codes.build_tables([0, 1, 2], [0, 0, 1])      # check 1 has degree 1
    # O(E): a million-edge graph builds in about a second (the reference's constructor needs hours there)
    import time
    bv, bc = codes.regular_ldpc(349_998, 3, 6, seed=1)
    t0 = time.time(); big = codes.build_tables(bv, bc); dt = time.time() - t0
    assert big["slot_edge"].size == bv.size and dt < 20
