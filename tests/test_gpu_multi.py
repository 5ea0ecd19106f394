"""cq_multi_* (reads sharded over the GPUs of one box behind the C ABI, one grouped NCCL reduce of
the counters) and cq_ilp_inputs (section 8f.3).

Sums of integers do not depend on the number of shards, so everything is compared for equality:
n GPUs == 1 GPU == oracle, including the per-read records each device writes into its slice of
the caller's buffers.  The 2-GPU tests need two devices; the 1-GPU tests run the same code path
with one shard (no NCCL)."""
import numpy as np
import pytest

import cammiq_b200 as cq
import oracle_lib as ol
from golden_util import golden_cases, load_case

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def assert_same(got, want, mode, per_read=True):
    keys = ["cnt_u", "cnt_d"] + (["read_class", "read_rid_a", "read_rid_b", "read_nleaf_u", "read_nleaf_d",
                                  "read_leaf_u", "read_leaf_d"] if per_read else [])
    for k in keys:
        assert np.array_equal(got[k], want[k]), k
    assert (int(got["nundet"]), int(got["nconf"]), int(got["n_invalid"])) == (
        int(want["nundet"]), int(want["nconf"]), int(want["n_invalid"]))
    if mode == cq.MODE_P:
        assert np.array_equal(got["rcount_u"], want["rcount_u"]) and np.array_equal(got["rcount_d"], want["rcount_d"])
    else:
        assert got["pairs"] == want["pairs"]


@pytest.mark.parametrize("n_gpus", [1, 2])
@pytest.mark.parametrize("case", golden_cases())
def test_multi_equals_single_and_oracle(case, n_gpus):
    if n_gpus > _gpus():
        pytest.skip("needs %d GPUs" % n_gpus)
    c = load_case(case)
    idx = cq.Index(c["iu"], c["id"])
    one = cq.Context(0).upload(idx, c["G"])
    multi = cq.MultiContext(n_gpus).upload(idx, c["G"])
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    for mode, omode in ((cq.MODE_P, ol.MODE_P), (cq.MODE_SC, ol.MODE_SC)):
        for packing in (0, 2):
            multi.set_host_packing(packing)
            want = one.query(mode, c["bases"], c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
            got = multi.query(mode, c["bases"], c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
            assert_same(got, want, mode)
            orc = ol.oracle_query(oi_u, oi_d, omode, c["G"], c["bases"], c["offsets"], c["lengths"], per_read=True, leaf_cap=128)
            assert_same(got, orc, mode)
            info = multi.info()
            assert info["n_gpus"] == n_gpus and sum(info["shard_reads"]) == len(c["lengths"])
            if n_gpus > 1:
                assert info["nccl_version"] > 0 and min(info["shard_reads"]) > 0
            # counters accumulate across calls until the reset, on every device
            again = multi.query(mode, c["bases"], c["offsets"], c["lengths"])
            assert np.array_equal(again["cnt_u"], 2 * want["cnt_u"]) and int(again["nundet"]) == 2 * int(want["nundet"])
            if mode == cq.MODE_P:
                assert np.array_equal(again["rcount_d"], 2 * want["rcount_d"])
            else:
                assert again["pairs"] == {k: 2 * v for k, v in want["pairs"].items()}
            one.reset()
            multi.reset()
    # reads the caller already holds packed
    pk, pl, _ = cq.pack_reads(c["bases"], c["offsets"], c["lengths"], threads=2)
    want = one.query_packed(cq.MODE_P, pk.reshape(-1), None, pl, stride=pk.shape[1])
    got = multi.query_packed(cq.MODE_P, pk.reshape(-1), None, pl, stride=pk.shape[1])
    assert_same(got, want, cq.MODE_P, per_read=False)
    multi.close()
    one.close()


def test_multi_rejects_bad_arguments():
    with pytest.raises(cq.CammiqError):
        cq.MultiContext(0)
    with pytest.raises(cq.CammiqError):
        cq.MultiContext(2, devices=[0, 0])
    with pytest.raises(cq.CammiqError):
        cq.MultiContext(1, devices=[63])
    m = cq.MultiContext(1)
    with pytest.raises(cq.CammiqError):
        m.query(cq.MODE_P, np.zeros(4, np.uint8), None, np.array([4], np.uint8), stride=4)   # no index resident
    m.close()


@pytest.mark.parametrize("case", ["cfg1_small", "adversarial_250"])
def test_ilp_inputs_on_the_device(case):
    """wcov = ucount * (rl - depth) * 1.0 / rl * pow(1 - erate, depth) per leaf (query.cpp:1157-1160,
    1171-1175), per-genome sums over map_sp (the COV coefficients of query.cpp:1196-1230) and
    per-genome rcount sums, against the formula evaluated on the host in numpy."""
    c = load_case(case)
    idx = cq.Index(c["iu"], c["id"])
    ctx = cq.Context(0).upload(idx, c["G"])
    res = ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    rl = int(c["lengths"].astype(np.uint64).sum() // len(c["lengths"]))       # query.cpp:1087
    erate = float(np.float32(0.01))                                           # the float option, widened
    got = ctx.ilp_inputs(erate, rl)
    G = c["G"]
    for table, keys, rc in ((cq.TABLE_U, ("wcov_u",), res["rcount_u"]), (cq.TABLE_D, ("wcov_d1", "wcov_d2"), res["rcount_d"])):
        lv = idx.leaves(table)
        depth = lv["depth"].astype(np.uint32)
        decay = np.power(1.0 - erate, depth.astype(np.float64))
        gsum, grc = np.zeros(G + 1), np.zeros(G + 1, dtype=np.uint64)
        for key, uc, ref in zip(keys, (lv["ucount1"], lv["ucount2"]), (lv["ref_id1"], lv["ref_id2"])):
            t = (uc.astype(np.uint32) * (np.uint32(rl) - depth)).astype(np.uint32)   # 32-bit unsigned, as in the reference
            want = t.astype(np.float64) * 1.0 / float(rl) * decay
            assert np.allclose(got[key], want, rtol=1e-12, atol=0.0), key
            np.add.at(gsum, ref, want)
            np.add.at(grc, ref, rc.astype(np.uint64))
        tag = "u" if table == cq.TABLE_U else "d"
        assert np.allclose(got["genome_wcov_" + tag][1:], gsum[1:], rtol=1e-10, atol=0.0)
        assert np.array_equal(got["genome_rcount_" + tag][1:], grc[1:])
    ctx.close()


@pytest.mark.parametrize("case", ["cfg1_small", "deep_h12"])
def test_ilp_inputs_against_the_reference_expressions(case, tmp_path):
    """`ref_harness ilp` evaluates the expressions of query.cpp:1087, 1157-1160, 1171-1175 on the
    reference's own nodes after the reference's own scan (the solver code around them is compiled out
    without CPLEX / Gurobi).  Per map_sp entry: rcount must be equal, wcov within 1e-12 relative."""
    import os
    import subprocess
    harness = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_harness")
    if not os.access(harness, os.X_OK):
        pytest.skip("oracle/_ref not built")
    c = load_case(case)
    out = str(tmp_path / "ilp.txt")
    subprocess.run([harness, "ilp", c["iu"], c["id"], c["map"], "0.01", out, c["fq"]], check=True, capture_output=True)
    rows = [l.split() for l in open(out)]
    rl = int(rows[0][1])
    idx = cq.Index(c["iu"], c["id"])
    ctx = cq.Context(0).upload(idx, c["G"])
    res = ctx.query(cq.MODE_P, c["bases"], c["offsets"], c["lengths"])
    got = ctx.ilp_inputs(float(np.float32(0.01)), rl)
    n = 0
    for tag, table, w1, w2, rc in (("U", cq.TABLE_U, got["wcov_u"], None, res["rcount_u"]),
                                   ("D", cq.TABLE_D, got["wcov_d1"], got["wcov_d2"], res["rcount_d"])):
        off, ids = idx.map_sp(table, c["G"])
        for r in rows[1:]:
            if r[0] != tag:
                continue
            g, k = int(r[1]), int(r[2])
            leaf = int(ids[int(off[g]) + k])
            assert int(rc[leaf]) == int(r[3])
            assert abs(w1[leaf] - float(r[4])) <= 1e-12 * max(abs(float(r[4])), 1e-300)
            if w2 is not None:
                assert abs(w2[leaf] - float(r[5])) <= 1e-12 * max(abs(float(r[5])), 1e-300)
            n += 1
    assert n == len(rows) - 1 and n > 100
    ctx.close()
