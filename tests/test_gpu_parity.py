"""GPU parity tests proper: the CUDA path through the C ABI (libcammiq_gpu.so) against
(a) the committed answers of the unmodified reference (tests/golden/) and (b) the oracle on
seeded inputs.  Exact integer equality everywhere."""
import os

import numpy as np
import pytest

import cammiq_b200 as cq
import oracle_lib as ol
import parity
from golden_util import golden_cases, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["ascii_over_pcie", "host_packed"])
def ctx(request):
    """Every test runs through both host->device paths of cq_query: ASCII reads decoded by the
    kernel, and reads packed to 2 bits per base by host threads before the copy."""
    c = cq.Context(0)
    c.set_host_packing(0 if request.param == "ascii_over_pcie" else 3)
    yield c
    c.close()


def run_gpu(ctx, c, mode, load_factor=0.0, filter_bytes=None, **kw):
    idx = cq.Index(c["iu"], c["id"], load_factor)
    if filter_bytes is not None:
        idx.set_filter_budget(filter_bytes)
        assert (idx.info.filter_bytes > 0) == (filter_bytes > 0)
    ctx.upload(idx, c["G"])
    return idx, ctx.query(cq.MODE_SC if mode == "sc" else cq.MODE_P, c["bases"], c["offsets"],
                          c["lengths"], **kw)


@pytest.mark.parametrize("mode", ["p", "sc"])
@pytest.mark.parametrize("case", golden_cases())
def test_gpu_matches_reference_dump(ctx, case, mode):
    c = load_case(case)
    dump = c["dump"][mode]
    idx, res = run_gpu(ctx, c, mode, per_read=True, leaf_cap=128)
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    assert res["n_invalid"] == 0
    parity.check_counters(res, dump["files"][0], oi_u, oi_d, c["G"], mode)
    parity.check_per_read(res, dump, oi_u, oi_d, mode)


@pytest.mark.parametrize("mode", ["p", "sc"])
@pytest.mark.parametrize("case", golden_cases())
def test_gpu_matches_oracle_every_read(ctx, case, mode):
    """All reads of the case, every per-read field, both load factors (0.95 forces multi-bucket
    probe chains)."""
    c = load_case(case)
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    m = ol.MODE_SC if mode == "sc" else ol.MODE_P
    want = ol.oracle_query(oi_u, oi_d, m, c["G"], c["bases"], c["offsets"], c["lengths"],
                           per_read=True, leaf_cap=128)
    # default filter; no filter (table probed in phase 1) with chained buckets; a tiny filter
    # with a high false-positive rate
    for lf, fb in ((0.0, None), (0.95, 0), (0.95, 8192)):
        idx, got = run_gpu(ctx, c, mode, load_factor=lf, filter_bytes=fb, per_read=True, leaf_cap=128)
        for k in ("nundet", "nconf", "n_invalid"):
            assert int(got[k]) == int(want[k]), (k, lf)
        for k in ("cnt_u", "cnt_d", "read_class", "read_rid_a", "read_rid_b", "read_nleaf_u",
                  "read_nleaf_d", "read_leaf_u", "read_leaf_d"):
            assert np.array_equal(got[k], want[k]), (k, lf)
        if mode == "p":
            assert np.array_equal(got["rcount_u"], want["rcount_u"])
            assert np.array_equal(got["rcount_d"], want["rcount_d"])
        else:
            assert got["pairs"] == want["pairs"]


@pytest.mark.parametrize("pairs", [1, 2, 4])
@pytest.mark.parametrize("case", golden_cases())
def test_sieve_regime_matches_oracle(ctx, case, pairs, monkeypatch):
    """The regime of indices too large for a selective filter (BASELINE configs[3]): the L2 filter
    as a sieve, passing positions load their bucket keys in phase 1.  Forced on the golden indices
    with 1, 2 and 4 bits per key (CAMMIQ_FILTER_FORCE_SIEVE) and a filter small enough that a good
    share of the positions pass."""
    c = load_case(case)
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    monkeypatch.setenv("CAMMIQ_FILTER_FORCE_SIEVE", str(pairs))
    for mode, m in (("p", ol.MODE_P), ("sc", ol.MODE_SC)):
        want = ol.oracle_query(oi_u, oi_d, m, c["G"], c["bases"], c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
        for lf in (0.0, 0.95):
            idx, got = run_gpu(ctx, c, mode, load_factor=lf, filter_bytes=8192, per_read=True, leaf_cap=128)
            for k in ("cnt_u", "cnt_d", "read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d",
                      "read_leaf_u", "read_leaf_d"):
                assert np.array_equal(got[k], want[k]), (k, lf, mode)
            assert (int(got["nundet"]), int(got["nconf"])) == (int(want["nundet"]), int(want["nconf"]))
            if mode == "p":
                assert np.array_equal(got["rcount_u"], want["rcount_u"]) and np.array_equal(got["rcount_d"], want["rcount_d"])
            else:
                assert got["pairs"] == want["pairs"]
            ctx.reset()


def test_counters_accumulate_and_reset(ctx):
    """Device counters live across calls until cq_reset, like the reference's counters until
    resetCounters (query.cpp:259-260, 1820-1840); splitting a file into batches is exact."""
    c = load_case("cfg1_small")
    idx = cq.Index(c["iu"], c["id"])
    ctx.upload(idx, c["G"])
    whole = ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    ctx.reset()
    n = len(c["lengths"])
    cuts = [0, 1, 2, n // 3, n // 3, n - 1, n]
    part = None
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = ctx.query(cq.MODE_P, c["bases"], c["offsets"][a:b], c["lengths"][a:b])
    for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d"):
        assert np.array_equal(whole[k], part[k]), k
    assert (whole["nundet"], whole["nconf"]) == (part["nundet"], part["nconf"])
    ctx.reset()
    zero = ctx.fetch(cq.MODE_P)
    assert zero["nundet"] == 0 and not zero["cnt_u"].any() and not zero["rcount_u"].any()


def test_invalid_and_short_reads(ctx):
    c = load_case("cfg1_small")
    idx = cq.Index(c["iu"], c["id"])
    ctx.upload(idx, c["G"])
    h = idx.hash_len
    reads = [c["reads"][0], c["reads"][0][:h - 1], c["reads"][1][:40] + b"R" + c["reads"][1][41:], b"",
             c["reads"][2].lower(), c["reads"][2][:h]]
    b, o, l = ol.pack_reads(reads)
    got = ctx.query(cq.MODE_P, b, o, l, per_read=True, leaf_cap=32)
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    want = ol.oracle_query(oi_u, oi_d, ol.MODE_P, c["G"], b, o, l, per_read=True, leaf_cap=32)
    assert got["n_invalid"] == want["n_invalid"] == 3
    for k in ("read_class", "read_rid_a", "read_rid_b", "cnt_u", "cnt_d", "rcount_u", "rcount_d"):
        assert np.array_equal(got[k], want[k]), k


def test_fixed_stride_reads_and_empty_batch(ctx):
    c = load_case("cfg1_small")
    idx = cq.Index(c["iu"], c["id"])
    ctx.upload(idx, c["G"])
    reads = [r for r in c["reads"] if len(r) == 100][:500]
    b, o, l = ol.pack_reads(reads)
    a = ctx.query(cq.MODE_P, b, o, l, per_read=True)
    ctx.reset()
    s = ctx.query(cq.MODE_P, b, None, l, stride=100, per_read=True)
    for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d", "read_class"):
        assert np.array_equal(a[k], s[k]), k
    ctx.reset()
    e = ctx.query(cq.MODE_P, np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint8))
    assert e["nundet"] == 0 and not e["cnt_u"].any()


def test_dense_index_spills_hit_list(ctx, tmp_path):
    """Dense synthetic index (a key at almost every position, branching buckets, a key in both
    tables): 250-base reads collect far more than the 64-entry shared hit list, exercising
    the global spill path, cross-chunk de-duplication and the trie descent."""
    from test_index_writer import make_dense_case
    seq, pu, pd, eu, ed, G = make_dense_case(tmp_path)
    oi_u, oi_d = ol.OracleIndex(pu), ol.OracleIndex(pd)
    rng = np.random.default_rng(9)
    reads = []
    for _ in range(60):
        ln = int(rng.integers(100, 256))
        st = int(rng.integers(0, len(seq) - ln))
        r = seq[st:st + ln]
        if rng.random() < 0.5:
            r = bytes(b"ACGT"[3 - b"ACGT".index(ch)] for ch in r[::-1])
        reads.append(r)
    for blk in range(4):                       # single-genome reads -> accepted, rcount path
        reads.append(seq[blk * 150 + 2:blk * 150 + 140])
        reads.append(seq[blk * 150 + 2:blk * 150 + 100] + seq[blk * 150 + 2:blk * 150 + 100])
    b, o, l = ol.pack_reads(reads)
    for mode, m in ((cq.MODE_P, ol.MODE_P), (cq.MODE_SC, ol.MODE_SC)):
        want = ol.oracle_query(oi_u, oi_d, m, G, b, o, l, per_read=True, leaf_cap=600)
        assert int((want["read_nleaf_u"] + want["read_nleaf_d"]).max()) > 100
        for lf, fb in ((0.0, None), (1.0, 0)):
            idx = cq.Index(pu, pd, lf)
            if fb is not None:
                idx.set_filter_budget(fb)
            ctx.upload(idx, G)
            got = ctx.query(mode, b, o, l, per_read=True, leaf_cap=600)
            for k in ("read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d",
                      "read_leaf_u", "read_leaf_d", "cnt_u", "cnt_d"):
                assert np.array_equal(got[k], want[k]), (k, lf)
            if mode == cq.MODE_P:
                assert np.array_equal(got["rcount_u"], want["rcount_u"])
                assert np.array_equal(got["rcount_d"], want["rcount_d"])
                assert want["rcount_u"].sum() > 0
            else:
                assert got["pairs"] == want["pairs"]


def test_chunked_pipeline_matches_single_launch_and_oracle(ctx, tmp_path):
    """cq_query streams reads in 2^20-read chunks through three rotating staging buffers (five
    chunks here, so buffers are reused while earlier chunks are still in flight); the result must
    equal the single-launch device-resident path, and per-read records around the chunk
    boundaries must equal the oracle (fixed-stride and offsets addressing, ragged lengths)."""
    from cammiq_b200 import synthlib as sl
    p = sl.params(seed=7, n_genomes=12, genome_len=200_000, cluster_size=3)
    sl.write_index(p, str(tmp_path))
    iu, idd = str(tmp_path / "index_u.bin1"), str(tmp_path / "index_d.bin2")
    idx = cq.Index(iu, idd)
    ctx.upload(idx, 12)
    n, rl = (1 << 22) + 12345, 100
    reads = sl.make_reads(p, 0, n, rl, 0.01)
    rng = np.random.default_rng(1)
    lengths = rng.integers(60, rl + 1, size=n).astype(np.uint8)
    lengths[::1000] = 10           # shorter than h
    reads[5::7777, 50] = ord("N")  # invalid byte
    flat = reads.reshape(-1)
    # (a) fixed stride, chunked host path vs device-resident single launch
    a = ctx.query(cq.MODE_P, flat, None, lengths, stride=rl, per_read=True)
    ctx.reset()
    ctx.stage(flat, None, lengths, stride=rl)
    ctx.query_staged(cq.MODE_P)
    b = ctx.fetch(cq.MODE_P)
    for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["nundet"], a["nconf"], a["n_invalid"]) == (b["nundet"], b["nconf"], b["n_invalid"])
    assert a["n_invalid"] == len(range(0, n, 1000)) + len(set(range(5, n, 7777)) - set(range(0, n, 1000)))
    # (b) offsets addressing (reads stored back to back, ragged), chunked
    offsets = np.zeros(n, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths.astype(np.uint64))[:-1]
    ragged = np.concatenate([reads[i, :lengths[i]] for i in range(0, n, 1)]) if n < 1000 else None
    if ragged is None:
        mask = np.arange(rl)[None, :] < lengths[:, None]
        ragged = reads[mask]
    ctx.reset()
    c2 = ctx.query(cq.MODE_P, ragged, offsets, lengths, per_read=True)
    for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d", "read_class", "read_rid_a", "read_rid_b"):
        assert np.array_equal(a[k], c2[k]), k
    # (c) oracle on windows around the chunk boundaries
    ou, od = ol.OracleIndex(iu), ol.OracleIndex(idd)
    for lo in (0, (1 << 20) - 1500, (1 << 21) - 1500, 3 * (1 << 20) - 1500, (1 << 22) - 1500, n - 3000):
        hi = min(lo + 3000, n)
        want = ol.oracle_query(ou, od, ol.MODE_P, 12, flat, np.arange(lo, hi, dtype=np.uint64) * rl,
                               lengths[lo:hi], per_read=True)
        for k in ("read_class", "read_rid_a", "read_rid_b"):
            assert np.array_equal(a[k][lo:hi], want[k]), (k, lo)
    assert (a["read_class"] >= 2).sum() > n // 10
