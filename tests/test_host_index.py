"""CPU tests of the host side of the C ABI: library loads and exports every declared symbol,
index decode == oracle decode, flattened lookup == oracle find64_p restatement.  No GPU."""
import os
import re

import numpy as np
import pytest

import cammiq_b200 as cq
import oracle_lib as ol
import synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")


def golden_cases():
    return sorted(d for d in os.listdir(GOLD) if os.path.isdir(os.path.join(GOLD, d)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "cammiq_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cq_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = cq.lib()
    for name in declared:
        assert hasattr(L, name), "libcammiq_gpu.so does not export " + name
    assert declared == set(cq.capi.SYMBOLS), declared ^ set(cq.capi.SYMBOLS)
    assert L.cq_abi_version() == cq.capi.ABI_VERSION == 3


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cq.CammiqError) as e:
        cq.Context(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_missing_index_file_is_an_error(tmp_path):
    with pytest.raises(cq.CammiqError) as e:
        cq.Index(str(tmp_path / "nope.bin1"), str(tmp_path / "nope.bin2"))
    assert e.value.code == -2


@pytest.mark.parametrize("case", golden_cases())
def test_decode_matches_oracle(case):
    d = os.path.join(GOLD, case)
    iu, idd = os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2")
    idx = cq.Index(iu, idd)
    for table, path in ((cq.TABLE_U, iu), (cq.TABLE_D, idd)):
        oi = ol.OracleIndex(path)
        lv = idx.leaves(table)
        assert idx.hash_len == oi.h
        assert len(lv["ref_id1"]) == oi.n_leaves
        for mine, ref in (("ref_id1", oi.ref1), ("ref_id2", oi.ref2), ("ucount1", oi.ucount1),
                          ("ucount2", oi.ucount2), ("depth", oi.depth)):
            assert np.array_equal(lv[mine], ref), (case, table, mine)
        G = int(max(oi.ref1.max(initial=0), oi.ref2.max(initial=0)))
        off, ids = idx.map_sp(table, G)
        ooff, oids = oi.map_sp(G)
        assert np.array_equal(off, ooff) and np.array_equal(ids, oids)


@pytest.mark.parametrize("threads", ["1", "5"])
def test_parallel_decode_of_many_buckets_matches_oracle(tmp_path, threads, monkeypatch):
    """The decoder's second pass fills 4096-bucket ranges in parallel from checkpoints of the
    first; an index with ~150 such ranges, a fifth of its keys deeper than h (tries that straddle
    range ends), against the oracle's sequential recursive decoder: every leaf field, the
    per-genome leaf lists in file order, and lookups that descend the tries."""
    from cammiq_b200 import synthlib as sl
    monkeypatch.setenv("CAMMIQ_DECODE_THREADS", threads)
    p = sl.params(seed=31, n_genomes=40, genome_len=400_000, cluster_size=4, permille_deep=200)
    sl.write_index(p, str(tmp_path))
    iu, idd = str(tmp_path / "index_u.bin1"), str(tmp_path / "index_d.bin2")
    idx = cq.Index(iu, idd)
    assert idx.info.n_buckets_u > 20 * 4096 and idx.info.n_nodes_u > 10000
    rng = np.random.default_rng(1)
    reads = sl.make_reads(p, 0, 300, 120, 0.0)
    for table, path in ((cq.TABLE_U, iu), (cq.TABLE_D, idd)):
        oi = ol.OracleIndex(path)
        lv = idx.leaves(table)
        for mine, ref in (("ref_id1", oi.ref1), ("ref_id2", oi.ref2), ("ucount1", oi.ucount1),
                          ("ucount2", oi.ucount2), ("depth", oi.depth)):
            assert np.array_equal(lv[mine], ref), (table, mine)
        off, ids = idx.map_sp(table, 40)
        ooff, oids = oi.map_sp(40)
        assert np.array_equal(off, ooff) and np.array_equal(ids, oids)
        h, hits = idx.hash_len, 0
        for r in reads:
            r = r.tobytes()
            for i in rng.integers(0, len(r) - h, 40):
                key = ol.lib().cqo_hash(r[i:i + h], h)
                want = oi.find(key, r[i + h:])
                assert idx.find_host(table, key, r[i + h:]) == want
                hits += want != ol.NONE
        assert hits > 20


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("load_factor,sieve", [(0.0, None), (0.95, None), (0.95, "1"), (0.0, "2")])
def test_flat_lookup_matches_oracle_find(case, load_factor, sieve, monkeypatch):
    """Every h-mer window of a read corpus (both strands): flattened table + path-compressed trie
    (behind the selective filter, or the 1- / 2-bit sieve of oversized indices) gives the same
    leaf as the oracle's find64_p (test pyramid item 2)."""
    d = os.path.join(GOLD, case)
    iu, idd = os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2")
    if sieve:
        monkeypatch.setenv("CAMMIQ_FILTER_FORCE_SIEVE", sieve)
    idx = cq.Index(iu, idd, load_factor)
    if sieve:
        idx.set_filter_budget(8192)
    ou, od = ol.OracleIndex(iu), ol.OracleIndex(idd)
    h = idx.hash_len
    reads = synth.read_fastq(os.path.join(d, "reads.fq"))[:300]
    hits = 0
    for r in reads:
        r = r[:255]
        for s in (r, synth.revcomp(np.frombuffer(r, dtype=np.uint8)).tobytes()):
            for i in range(0, len(s) - h + 1):
                hv = ol.lib().cqo_hash(s[i:i + h], h)
                cand = s[i + h:]
                for table, oi in ((cq.TABLE_U, ou), (cq.TABLE_D, od)):
                    want = oi.find(hv, cand)
                    got = idx.find_host(table, hv, cand)
                    assert got == want, (case, table, i)
                    hits += want != ol.NONE
    assert hits > 0


def test_empty_index_pair(tmp_path):
    """An empty but valid index: INT = 10 x 0xFF, AUX = flag|64, h."""
    for name, first in (("e.bin1", 0x40), ("e.bin2", 0xC0)):
        (tmp_path / name).write_bytes(b"\xff" * 10)
        (tmp_path / (name + ".aux")).write_bytes(bytes([first, 26]) + b"\xff" * 9)
    idx = cq.Index(str(tmp_path / "e.bin1"), str(tmp_path / "e.bin2"))
    assert (idx.hash_len, idx.n_leaves_u, idx.n_leaves_d, idx.info.n_keys) == (26, 0, 0, 0)
    assert idx.find_host(cq.TABLE_U, 12345, b"ACGT") == ol.NONE


def test_mismatched_hash_lengths_rejected(tmp_path):
    (tmp_path / "a.bin1").write_bytes(b"\xff" * 10)
    (tmp_path / "a.bin1.aux").write_bytes(bytes([0x40, 26]) + b"\xff" * 9)
    (tmp_path / "a.bin2").write_bytes(b"\xff" * 10)
    (tmp_path / "a.bin2.aux").write_bytes(bytes([0xC0, 20]) + b"\xff" * 9)
    with pytest.raises(cq.CammiqError) as e:
        cq.Index(str(tmp_path / "a.bin1"), str(tmp_path / "a.bin2"))
    assert e.value.code == -3


def test_truncated_index_rejected(tmp_path):
    d = os.path.join(GOLD, golden_cases()[0])
    data = open(os.path.join(d, "index_u.bin1"), "rb").read()
    (tmp_path / "t.bin1").write_bytes(data[:len(data) // 2])
    (tmp_path / "t.bin1.aux").write_bytes(open(os.path.join(d, "index_u.bin1.aux"), "rb").read())
    (tmp_path / "t.bin2").write_bytes(open(os.path.join(d, "index_d.bin2"), "rb").read())
    (tmp_path / "t.bin2.aux").write_bytes(open(os.path.join(d, "index_d.bin2.aux"), "rb").read())
    with pytest.raises(cq.CammiqError) as e:
        cq.Index(str(tmp_path / "t.bin1"), str(tmp_path / "t.bin2"))
    assert e.value.code == -3
