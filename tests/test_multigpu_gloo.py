"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups run the same sharding and
combine code bench.py runs over NCCL.  Each rank's scan result is produced by the oracle on its
shard (stand-in for the per-GPU kernel result); the combined counters on rank 0 must equal the
oracle on all reads -- exact integer equality."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, world, port, case_name, mode, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, REPO)
    import oracle_lib as ol
    from cammiq_b200 import multigpu
    from golden_util import load_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = load_case(case_name)
    G = c["G"]
    oi_u, oi_d = ol.OracleIndex(c["iu"]), ol.OracleIndex(c["id"])
    n = len(c["lengths"])
    lo, hi = multigpu.shard_range(n, rank, world)
    m = ol.MODE_SC if mode == "sc" else ol.MODE_P
    r = ol.oracle_query(oi_u, oi_d, m, G, c["bases"], c["offsets"][lo:hi], c["lengths"][lo:hi])
    # the same layout the device exposes: uint64 counter block viewed as int64, uint32 rcount as int32
    block = np.concatenate([r["cnt_u"], r["cnt_d"],
                            np.array([r["nundet"], r["nconf"], r["n_invalid"], 0], dtype=np.uint64)])
    counts = torch.from_numpy(block.view(np.int64).copy())
    rc_u = torch.from_numpy(r["rcount_u"].view(np.int32).copy())
    rc_d = torch.from_numpy(r["rcount_d"].view(np.int32).copy())
    multigpu.combine_counters(counts, rc_u if mode == "p" else None, rc_d if mode == "p" else None)
    pairs = multigpu.gather_pair_maps(r["pairs"])
    if rank == 0:
        want = ol.oracle_query(oi_u, oi_d, m, G, c["bases"], c["offsets"], c["lengths"])
        got = multigpu.unpack_counts(counts, G)
        ok = (got["cnt_u"] == [int(x) for x in want["cnt_u"]] and got["cnt_d"] == [int(x) for x in want["cnt_d"]]
              and got["nundet"] == want["nundet"] and got["nconf"] == want["nconf"])
        if mode == "p":
            ok = ok and np.array_equal(rc_u.numpy().view(np.uint32), want["rcount_u"])
            ok = ok and np.array_equal(rc_d.numpy().view(np.uint32), want["rcount_d"])
        else:
            ok = ok and pairs == want["pairs"]
        open(out_path, "w").write("ok" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "p"), (2, "sc"), (3, "p")])
def test_sharded_combine_equals_single_shard(tmp_path, world, mode):
    out = str(tmp_path / "result.txt")
    port = 29600 + (os.getpid() + world * 7 + len(mode)) % 300
    mp.spawn(_worker, args=(world, port, "cfg1_small", mode, out), nprocs=world, join=True)
    assert open(out).read() == "ok"


def test_shard_ranges_partition():
    from cammiq_b200 import multigpu
    for n in (0, 1, 7, 1000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            edges = [multigpu.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
