"""Comparison helpers shared by the oracle-vs-reference and GPU-vs-oracle parity tests."""
import numpy as np


def check_leaf_tables(dump, oi_u, oi_d):
    """Index decode parity: per-genome map_sp lists (order!) and every leaf field."""
    G = dump["g"]
    for key, oi in (("leaf_u", oi_u), ("leaf_d", oi_d)):
        canon = oi.canonical_ids()
        off, ids = oi.map_sp(G)
        for rid in range(1, G + 1):
            mine = ids[int(off[rid]):int(off[rid + 1])]
            ref = dump[key].get(rid, [])
            assert len(mine) == len(ref), (key, rid)
            for l, t in zip(mine, ref):
                l = int(l)
                assert (int(canon[l]), int(oi.ref1[l]), int(oi.ref2[l]), int(oi.depth[l]),
                        int(oi.ucount1[l]), int(oi.ucount2[l])) == t, (key, rid, l)


def check_counters(res, fdump, oi_u, oi_d, G, mode):
    """Aggregate parity for one FASTQ file: nundet, nconf, cnt_u/d, per-leaf rcount (mode p)
    in map_sp order, pair map (mode sc).  Exact integer equality."""
    assert int(res["nundet"]) == fdump["nundet"]
    assert int(res["nconf"]) == fdump["nconf"]
    assert [int(x) for x in res["cnt_u"][1:G + 1]] == fdump["cu"]
    assert [int(x) for x in res["cnt_d"][1:G + 1]] == fdump["cd"]
    if mode == "sc":
        assert res["pairs"] == fdump["pairs"]
        return
    for key, oi, rc in (("rcu", oi_u, res["rcount_u"]), ("rcd", oi_d, res["rcount_d"])):
        off, ids = oi.map_sp(G)
        for rid in range(1, G + 1):
            mine = [int(rc[int(l)]) for l in ids[int(off[rid]):int(off[rid + 1])]]
            assert mine == fdump[key].get(rid, []), (key, rid)


def expected_increments(cls, a, b, mode):
    """(nundet, nconf, u-list, d-list) implied by a per-read record."""
    if cls == 0:
        return 1, 0, [], []
    if cls == 1:
        return 0, 1, [], []
    if cls == 2:
        return 0, 0, [a], []
    if cls == 3:
        return 0, 0, [], sorted([a, b])
    if cls == 4:
        return 0, 0, [a], [a]
    if cls == 5:
        return (0, 0, [a], [a]) if mode == "sc" else (0, 0, [], [a])
    raise AssertionError(cls)


def check_per_read(res, dump, oi_u, oi_d, mode):
    """Per-read parity against the reference's own single-read decisions and leaf sets."""
    cu, cd = oi_u.canonical_ids(), oi_d.canonical_ids()
    for rec in dump["reads"]:
        i = rec["idx"]
        exp = expected_increments(int(res["read_class"][i]), int(res["read_rid_a"][i]),
                                  int(res["read_rid_b"][i]), mode)
        assert exp == (rec["nundet"], rec["nconf"], rec["u"], rec["d"]), (i, exp, rec)
        if "read_nleaf_u" in res:
            nu, nd = int(res["read_nleaf_u"][i]), int(res["read_nleaf_d"][i])
            lu = sorted(int(cu[int(l)]) for l in res["read_leaf_u"][i][:nu])
            ld = sorted(int(cd[int(l)]) for l in res["read_leaf_d"][i][:nd])
            assert lu == rec["lu"], (i, lu, rec["lu"])
            assert ld == rec["ld"], (i, ld, rec["ld"])
