"""The test-side index writer produces files both decoders read identically, including
branching buckets, deep tries and keys shared by both tables."""
import numpy as np

import cammiq_b200 as cq
import oracle_lib as ol
from index_writer import write_index


def make_dense_case(tmp_path, h=12, G=4, glen=600, seed=5):
    rng = np.random.default_rng(seed)
    seq = bytes(b"ACGT"[i] for i in rng.integers(0, 4, glen))
    seen, eu, ed = set(), [], []
    for pos in range(0, glen - h - 6):
        ln = h + (pos % 5)           # lengths h..h+4 -> tries of depth 0..4, siblings branch
        key = seq[pos:pos + ln]
        if any(key[:k] in seen for k in range(h, ln + 1)) or any(s.startswith(key) for s in seen):
            continue
        seen.add(key)
        block = pos // 150            # one genome per 150-base block
        if pos % 3 == 2:
            ed.append((key, 1 + block % G, 1 + (block + 1) % G, 1 + pos % 7, 2))
        else:
            eu.append((key, 1 + block % G, 0, 1 + pos % 5, 0))
    # a key present in BOTH tables (separate leaf id spaces, both counted)
    ed.append((eu[0][0], 1, 2, 3, 4))
    pu, pd = str(tmp_path / "dense_u.bin1"), str(tmp_path / "dense_d.bin2")
    write_index(pu, h, eu, False)
    write_index(pd, h, ed, True)
    return seq, pu, pd, eu, ed, G


def test_writer_roundtrip_both_decoders(tmp_path):
    seq, pu, pd, eu, ed, G = make_dense_case(tmp_path)
    idx = cq.Index(pu, pd)
    for table, path, entries in ((cq.TABLE_U, pu, eu), (cq.TABLE_D, pd, ed)):
        oi = ol.OracleIndex(path)
        lv = idx.leaves(table)
        assert oi.n_leaves == len(entries) == len(lv["ref_id1"])
        assert np.array_equal(lv["ref_id1"], oi.ref1) and np.array_equal(lv["depth"], oi.depth)
        h = oi.h
        for key, r1, r2, c1, c2 in entries:
            hv = ol.lib().cqo_hash(key[:h], h)
            want = oi.find(hv, key[h:] + b"ACGT")
            assert want != ol.NONE and int(oi.ref1[want]) == r1 and int(oi.depth[want]) == len(key)
            assert idx.find_host(table, hv, key[h:] + b"ACGT") == want
            # one base short of the key: no match
            if len(key) > h:
                assert idx.find_host(table, hv, key[h:-1]) == ol.NONE == oi.find(hv, key[h:-1])
    assert idx.info.n_nodes_u > 0 and idx.info.n_keys < idx.info.n_buckets_u + idx.info.n_buckets_d
