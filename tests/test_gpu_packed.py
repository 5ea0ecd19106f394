"""cq_query_packed / cq_pack_reads: reads packed to 2 bits per base by the
caller must give exactly the results of the ASCII entry point, which the other GPU tests pin
to the reference dumps and the oracle."""
import numpy as np
import pytest

import cammiq_b200 as cq
import oracle_lib as ol
from golden_util import golden_cases, load_case

pytestmark = pytest.mark.gpu

KEYS = ("cnt_u", "cnt_d", "read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d",
        "read_leaf_u", "read_leaf_d")


@pytest.fixture(scope="module")
def ctx():
    c = cq.Context(0)
    yield c
    c.close()


def same(a, b, mode):
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert (a["nundet"], a["nconf"], a["n_invalid"]) == (b["nundet"], b["nconf"], b["n_invalid"])
    if mode == cq.MODE_P:
        assert np.array_equal(a["rcount_u"], b["rcount_u"]) and np.array_equal(a["rcount_d"], b["rcount_d"])
    else:
        assert a["pairs"] == b["pairs"]


@pytest.mark.parametrize("mode", [cq.MODE_P, cq.MODE_SC])
@pytest.mark.parametrize("case", golden_cases())
def test_prepacked_reads_match_ascii_and_oracle(ctx, case, mode):
    c = load_case(case)
    ctx.upload(cq.Index(c["iu"], c["id"]), c["G"])
    ou, od = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    want = ol.oracle_query(ou, od, ol.MODE_SC if mode == cq.MODE_SC else ol.MODE_P, c["G"], c["bases"],
                           c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
    ctx.set_host_packing(0)
    ascii_res = ctx.query(mode, c["bases"], c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
    ctx.reset()
    same(ascii_res, want, mode)
    packed, plen, bad = cq.pack_reads(c["bases"], c["offsets"], c["lengths"], threads=2)
    assert bad == 0
    # fixed stride
    got = ctx.query_packed(mode, packed.reshape(-1), None, plen, stride=packed.shape[1], per_read=True, leaf_cap=128)
    ctx.reset()
    same(got, want, mode)
    # back-to-back with explicit offsets, in shuffled storage order
    nbytes = (plen.astype(np.int64) + 3) // 4
    order = np.random.default_rng(5).permutation(len(plen))
    offs = np.zeros(len(plen), dtype=np.uint64)
    offs[order] = np.concatenate([[0], np.cumsum(nbytes[order])[:-1]]).astype(np.uint64)
    blob = np.zeros(int(nbytes.sum()) + 1, dtype=np.uint8)
    for i in range(len(plen)):
        blob[int(offs[i]):int(offs[i]) + int(nbytes[i])] = packed[i, :nbytes[i]]
    got = ctx.query_packed(mode, blob, offs, plen, per_read=True, leaf_cap=128)
    ctx.reset()
    same(got, want, mode)


def test_invalid_and_short_reads_through_the_packer(ctx):
    c = load_case("cfg1_small")
    ctx.upload(cq.Index(c["iu"], c["id"]), c["G"])
    reads = [bytes(c["bases"][int(o):int(o) + int(l)]) for o, l in zip(c["offsets"][:400], c["lengths"][:400])]
    reads[3] = reads[3][:40] + b"N" + reads[3][41:]
    reads[77] = b"ACGT"                      # shorter than h
    reads[200] = reads[200][:-1] + b"x"      # bad last base
    reads[399] = b""
    b, o, l = ol.pack_reads(reads)
    ou, od = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    want = ol.oracle_query(ou, od, ol.MODE_P, c["G"], b, o, l, per_read=True, leaf_cap=64)
    for threads in (0, 1, 5):
        ctx.set_host_packing(threads)
        got = ctx.query(cq.MODE_P, b, o, l, per_read=True, leaf_cap=64)
        ctx.reset()
        same(got, want, cq.MODE_P)
        t = ctx.timing()
        assert t["host_pack_threads"] == threads
        assert t["h2d_bytes"] > 0
    assert want["n_invalid"] == 4
    packed, plen, bad = cq.pack_reads(b, o, l, threads=3)
    assert bad == 2 and plen[3] == 0 and plen[200] == 0 and plen[77] == 4
    got = ctx.query_packed(cq.MODE_P, packed.reshape(-1), None, plen, stride=packed.shape[1], per_read=True, leaf_cap=64)
    ctx.reset()
    same(got, want, cq.MODE_P)


def test_host_packing_moves_a_quarter_of_the_bytes(ctx):
    c = load_case("cfg1_small")
    ctx.upload(cq.Index(c["iu"], c["id"]), c["G"])
    ctx.set_host_packing(0)
    ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    plain = ctx.timing()["h2d_bytes"]
    ctx.reset()
    ctx.set_host_packing(2)
    ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    packed = ctx.timing()["h2d_bytes"]
    ctx.reset()
    n, total = len(c["lengths"]), int(c["lengths"].astype(np.int64).sum())
    assert plain >= total + 9 * n
    assert packed <= total // 4 + n + 5 * n + 64   # codes + lengths + 32-bit offsets


def test_hybrid_transfer_equals_single_call(ctx, monkeypatch):
    """cq_query from a page-locked buffer sends every n-th chunk over PCIe as ASCII while the host
    threads pack the others (CAMMIQ_DIRECT_EVERY): 3.3M reads = 4 chunks through both routes and
    the rotating stage buffers; the totals must be those of the reads queried once, times their
    multiplicity, whatever the mix."""
    import ctypes as C
    c = load_case("long150_h20")
    ctx.upload(cq.Index(c["iu"], c["id"]), c["G"])
    ctx.set_host_packing(0)
    once = ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    ctx.reset()
    reps = 4700                                               # 700 reads x 4700 = 3.29M reads
    n1, nb = len(c["lengths"]), len(c["bases"])
    p = C.c_void_p()
    cq.capi._check(cq.lib().cq_host_alloc(nb * reps, C.byref(p)))
    try:
        bases = np.frombuffer((C.c_char * (nb * reps)).from_address(p.value), dtype=np.uint8)
        bases[:] = np.tile(c["bases"], reps)
        lengths = np.tile(c["lengths"], reps)
        offsets = (np.tile(c["offsets"], reps) + np.repeat(np.arange(reps, dtype=np.uint64) * np.uint64(nb), n1)).astype(np.uint64)
        for every, threads in (("0", 3), ("2", 3), ("3", 2), ("1", 2)):
            monkeypatch.setenv("CAMMIQ_DIRECT_EVERY", every)
            ctx.set_host_packing(threads)
            got = ctx.query(cq.MODE_P, bases, offsets, lengths)
            ctx.reset()
            assert np.array_equal(got["cnt_u"], once["cnt_u"] * reps), every
            assert np.array_equal(got["cnt_d"], once["cnt_d"] * reps), every
            assert np.array_equal(got["rcount_u"], once["rcount_u"] * reps) and np.array_equal(got["rcount_d"], once["rcount_d"] * reps)
            assert (int(got["nundet"]), int(got["nconf"])) == (int(once["nundet"]) * reps, int(once["nconf"]) * reps)
    finally:
        ctx.set_host_packing(-1)
        cq.lib().cq_host_free(p)
