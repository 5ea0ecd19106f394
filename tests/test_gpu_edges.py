"""Edge cases of the CUDA path against the oracle: extreme hash lengths, more genomes than the
shared-memory counters hold (global-atomic path), sparse / non-monotonic read offsets (the pack
kernel gathers every read from wherever it lies), maximum read length, both tables empty, and the
size-independent properties of the scan at the benchmark's full size."""
import os

import numpy as np
import pytest

import cammiq_b200 as cq
import oracle_lib as ol
from index_writer import write_index

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGT", b"TGCA")


@pytest.fixture(scope="module", params=["ascii_over_pcie", "host_packed"])
def ctx(request):
    """Every test runs through both host->device paths of cq_query: ASCII reads decoded by the
    kernel, and reads packed to 2 bits per base by host threads before the copy."""
    c = cq.Context(0)
    c.set_host_packing(0 if request.param == "ascii_over_pcie" else 3)
    yield c
    c.close()


def random_seq(rng, n):
    return bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))


def build_case(tmp_path, h, G, seq, stride=3, deep=4, tag="x"):
    """Keys taken from `seq` every `stride` bases, lengths h..h+deep, ids spread over 1..G."""
    seen, eu, ed = set(), [], []
    for k, pos in enumerate(range(0, len(seq) - h - deep - 1, stride)):
        key = seq[pos:pos + h + (k % (deep + 1))]
        if any(key[:j] in seen for j in range(h, len(key) + 1)) or any(s.startswith(key) for s in seen):
            continue
        seen.add(key)
        a = 1 + ((pos // 300) * 7919) % G    # one genome per 300-base block: reads inside a block are assignable
        if k % 3 == 0:
            b = 1 + (a * 31 + 5) % G
            ed.append((key, a, b if b != a else 1 + a % G, 1 + k % 3, 2))
        else:
            eu.append((key, a, 0, 1 + k % 4, 0))
    pu, pd = str(tmp_path / (tag + "_u.bin1")), str(tmp_path / (tag + "_d.bin2"))
    write_index(pu, h, eu, False)
    write_index(pd, h, ed, True)
    return pu, pd


def compare(ctx, pu, pd, G, reads, modes=("p", "sc"), offsets=None, bases=None, lengths=None, leaf_cap=64):
    oi_u, oi_d = ol.OracleIndex(pu), ol.OracleIndex(pd)
    if bases is None:
        bases, offsets, lengths = ol.pack_reads(reads)
    idx = cq.Index(pu, pd)
    ctx.upload(idx, G)
    for mode in modes:
        m = ol.MODE_SC if mode == "sc" else ol.MODE_P
        want = ol.oracle_query(oi_u, oi_d, m, G, bases, offsets, lengths, per_read=True, leaf_cap=leaf_cap)
        got = ctx.query(cq.MODE_SC if mode == "sc" else cq.MODE_P, bases, offsets, lengths, per_read=True, leaf_cap=leaf_cap)
        for k in ("read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d", "read_leaf_u",
                  "read_leaf_d", "cnt_u", "cnt_d"):
            assert np.array_equal(got[k], want[k]), (k, mode)
        assert (int(got["nundet"]), int(got["nconf"]), int(got["n_invalid"])) == (want["nundet"], want["nconf"], want["n_invalid"])
        if mode == "p":
            assert np.array_equal(got["rcount_u"], want["rcount_u"]) and np.array_equal(got["rcount_d"], want["rcount_d"])
        else:
            assert got["pairs"] == want["pairs"]
        ctx.reset()
    return want


def sample_reads(rng, seq, n, lo, hi, rc_prob=0.5, err=0.0):
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi + 1))
        st = int(rng.integers(0, len(seq) - ln))
        r = bytearray(seq[st:st + ln])
        for i in range(ln):
            if err and rng.random() < err:
                r[i] = b"ACGT"[int(rng.integers(0, 4))]
        r = bytes(r)
        if rng.random() < rc_prob:
            r = r.translate(COMP)[::-1]
        out.append(r)
    return out


@pytest.mark.parametrize("h", [5, 16, 31])
def test_extreme_hash_lengths(ctx, tmp_path, h):
    rng = np.random.default_rng(100 + h)
    seq = random_seq(rng, 3000 if h > 5 else 400)
    pu, pd = build_case(tmp_path, h, 7, seq, stride=2 if h > 5 else 1, tag="h%d" % h)
    reads = sample_reads(rng, seq, 300, h, 255, err=0.01)
    reads += [seq[10:10 + h], seq[20:20 + h - 1] if h > 1 else b"A", seq[:255]]
    want = compare(ctx, pu, pd, 7, reads, leaf_cap=600)
    assert (want["read_class"] >= 2).sum() > 0


@pytest.mark.parametrize("h", [6, 8])
def test_palindromic_hmers(ctx, tmp_path, h):
    """An h-mer equal to its own reverse complement (even h only) is found by BOTH strands of a
    window; phase 1 keeps no flag for it, phase 2 serves the reverse strand from the forward
    candidate.  The sequence is seeded with palindromes so that many keys are palindromic."""
    rng = np.random.default_rng(600 + h)
    parts = []
    for _ in range(250):
        half = random_seq(rng, h // 2)
        parts.append(half + half.translate(COMP)[::-1])          # a palindromic h-mer
        parts.append(random_seq(rng, int(rng.integers(3, 9))))
    seq = b"".join(parts)
    pu, pd = build_case(tmp_path, h, 5, seq, stride=1, deep=3, tag="pal%d" % h)
    reads = sample_reads(rng, seq, 400, h, 200, err=0.0)
    reads += [p for p in parts[0:40:2]]                           # the palindromes alone
    want = compare(ctx, pu, pd, 5, reads, leaf_cap=600)
    assert (want["read_nleaf_u"] + want["read_nleaf_d"]).max() > 4


def test_more_genomes_than_shared_counters(ctx, tmp_path):
    """G = 9000 > 1023: genome counters fall back to global 64-bit atomics."""
    rng = np.random.default_rng(7)
    seq = random_seq(rng, 6000)
    G = 9000
    pu, pd = build_case(tmp_path, 20, G, seq, stride=5, tag="bigG")
    reads = sample_reads(rng, seq, 2000, 60, 150, err=0.005)
    want = compare(ctx, pu, pd, G, reads)
    assert want["cnt_u"].sum() > 0 and want["cnt_d"].sum() > 0 and int(np.count_nonzero(want["cnt_u"])) > 10


def test_sparse_and_shuffled_offsets_use_direct_loads(ctx, tmp_path):
    """Reads scattered over a buffer with gaps and in shuffled order, at every byte alignment:
    the pack kernel assembles each 16-base word from the aligned words it touches."""
    rng = np.random.default_rng(9)
    seq = random_seq(rng, 5000)
    pu, pd = build_case(tmp_path, 18, 5, seq, stride=4, tag="sparse")
    reads = sample_reads(rng, seq, 700, 50, 120, err=0.01)
    gap = 4096
    order = rng.permutation(len(reads))
    bases = np.frombuffer(random_seq(rng, 64) * ((gap * len(reads)) // 64 + 8), dtype=np.uint8).copy()
    offsets = np.zeros(len(reads), dtype=np.uint64)
    lengths = np.zeros(len(reads), dtype=np.uint8)
    for slot, i in enumerate(order):
        o = slot * gap + int(rng.integers(0, 3000))
        bases[o:o + len(reads[i])] = np.frombuffer(reads[i], dtype=np.uint8)
        offsets[i] = o
        lengths[i] = len(reads[i])
    compare(ctx, pu, pd, 5, None, bases=bases, offsets=offsets, lengths=lengths)
    # and overlapping reads (every read starts 1 byte after the previous one)
    n = 1500
    off2 = np.arange(n, dtype=np.uint64) * 1
    len2 = np.full(n, 100, dtype=np.uint8)
    b2 = np.frombuffer(seq[:n + 200], dtype=np.uint8).copy()
    compare(ctx, pu, pd, 5, None, bases=b2, offsets=off2, lengths=len2)


def test_two_accumulator_sets(ctx, tmp_path):
    """cq_swap_accumulators: two batches scanned into the two sets stay apart (each set holds exactly its
    own batch, counters and per-leaf rcount), resets touch only the current set, and the arrays the swap
    reports are those of the set that was current."""
    rng = np.random.default_rng(31)
    seq = random_seq(rng, 5000)
    G = 6
    pu, pd = build_case(tmp_path, 18, G, seq, stride=3, tag="swap")
    oi_u, oi_d = ol.OracleIndex(pu), ol.OracleIndex(pd)
    batches = [sample_reads(rng, seq, 500, 50, 150, err=0.01), sample_reads(rng, seq, 300, 40, 200, err=0.0)]
    packed = [ol.pack_reads(b) for b in batches]
    want = [ol.oracle_query(oi_u, oi_d, ol.MODE_P, G, *pk) for pk in packed]
    assert not np.array_equal(want[0]["rcount_u"], want[1]["rcount_u"])
    ctx.upload(cq.Index(pu, pd), G)

    def check(got, w):
        for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d"):
            assert np.array_equal(got[k], w[k]), k
        assert (int(got["nundet"]), int(got["nconf"])) == (w["nundet"], w["nconf"])

    first = ctx.device_counters()
    check(ctx.query(cq.MODE_P, *packed[0]), want[0])            # set A <- batch 0
    prev = ctx.swap_accumulators()                                # B current (fresh: zeroed)
    assert prev[1].ptr == first.d_rcount_u and ctx.device_counters().d_rcount_u != first.d_rcount_u
    check(ctx.query(cq.MODE_P, *packed[1]), want[1])            # set B <- batch 1 only
    ctx.swap_accumulators()                                       # A current again, untouched by batch 1
    assert ctx.device_counters().d_rcount_u == first.d_rcount_u
    check(ctx.fetch(cq.MODE_P), want[0])
    ctx.reset()                                                   # zeroes A, not B
    ctx.swap_accumulators()
    check(ctx.fetch(cq.MODE_P), want[1])
    ctx.reset()
    ctx.swap_accumulators()
    got = ctx.fetch(cq.MODE_P)
    assert int(got["rcount_u"].sum()) == 0 and int(got["cnt_u"].sum()) == 0


def test_empty_index_pair(ctx, tmp_path):
    for name, first in (("e.bin1", 0x40), ("e.bin2", 0xC0)):
        (tmp_path / name).write_bytes(b"\xff" * 10)
        (tmp_path / (name + ".aux")).write_bytes(bytes([first, 26]) + b"\xff" * 9)
    rng = np.random.default_rng(1)
    reads = [random_seq(rng, 100) for _ in range(100)]
    want = compare(ctx, str(tmp_path / "e.bin1"), str(tmp_path / "e.bin2"), 3, reads)
    assert want["nundet"] == 100


def test_full_size_properties(ctx, tmp_path):
    """BASELINE configs[1] size (500 genomes, 10M x 100bp reads): properties that need no oracle.
    (1) strand symmetry: reverse-complementing every read leaves every counter unchanged;
    (2) order invariance: a permutation of the reads leaves every counter unchanged;
    (3) additivity: two halves accumulated == the whole;
    (4) conservation: unlabeled + conflict + accepted reads == n, accepted classes bound the sums;
    (5) the first 20k reads equal the oracle."""
    from cammiq_b200 import synthlib as sl
    p = sl.params(seed=2, n_genomes=500, genome_len=3_000_000, cluster_size=4)
    sl.write_index(p, str(tmp_path))
    iu, idd = str(tmp_path / "index_u.bin1"), str(tmp_path / "index_d.bin2")
    idx = cq.Index(iu, idd)
    ctx.upload(idx, 500)
    n, rl = 10_000_000, 100
    reads = sl.make_reads(p, 0, n, rl, 0.01)
    lengths = np.full(n, rl, dtype=np.uint8)
    keys = ("cnt_u", "cnt_d", "rcount_u", "rcount_d")
    a = ctx.query(cq.MODE_P, reads.reshape(-1), None, lengths, stride=rl, per_read=True)
    ctx.reset()
    # (1)
    lut = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        lut[x] = y
    rc = np.ascontiguousarray(lut[reads[:, ::-1]])
    b = ctx.query(cq.MODE_P, rc.reshape(-1), None, lengths, stride=rl, per_read=True)
    ctx.reset()
    for k in keys + ("read_class", "read_rid_a", "read_rid_b"):
        assert np.array_equal(a[k], b[k]), k
    # (2)
    perm = np.random.default_rng(0).permutation(n)
    offs = perm.astype(np.uint64) * rl          # read i of the query = read perm[i] of the buffer
    c = ctx.query(cq.MODE_P, reads.reshape(-1), offs, lengths, per_read=True)
    ctx.reset()
    for k in keys:
        assert np.array_equal(a[k], c[k]), k
    assert np.array_equal(a["read_class"][perm], c["read_class"])
    # (3)
    half = n // 2 + 12345
    ctx.query(cq.MODE_P, reads[:half].reshape(-1), None, lengths[:half], stride=rl)
    d = ctx.query(cq.MODE_P, reads[half:].reshape(-1), None, lengths[half:], stride=rl)
    ctx.reset()
    for k in keys:
        assert np.array_equal(a[k], d[k]), k
    assert (a["nundet"], a["nconf"]) == (d["nundet"], d["nconf"])
    # (4)
    cls = a["read_class"]
    assert int((cls == 0).sum()) == a["nundet"] and int((cls == 1).sum()) == a["nconf"]
    assert int(a["cnt_u"].sum()) == int(((cls == 2) | (cls == 4)).sum())
    assert int(a["cnt_d"].sum()) == int(2 * (cls == 3).sum() + (cls == 4).sum() + (cls == 5).sum())
    assert int(a["rcount_u"].sum() + a["rcount_d"].sum()) >= int((cls >= 2).sum())
    # (5)
    m = 20_000
    oi_u, oi_d = ol.OracleIndex(iu), ol.OracleIndex(idd)
    want = ol.oracle_query(oi_u, oi_d, ol.MODE_P, 500, reads[:m].reshape(-1), np.arange(m, dtype=np.uint64) * rl,
                           lengths[:m], per_read=True)
    for k in ("read_class", "read_rid_a", "read_rid_b"):
        assert np.array_equal(a[k][:m], want[k]), k
