#!/usr/bin/env python
"""A/B timing of scan-kernel builds on one workload: resident reads, CUDA-event scan time.

    CAMMIQ_LIB=cammiq_b200/variants/libcammiq_gpu_s8b3.so python tools/kernel_ab.py [--workload cfg2]
        [--reads N] [--filter-mb M] [--packed] [--mode p|sc] [--iters K]

Prints one JSON line: scan ms (mean / min over K launches), registers, grid, a checksum of the
counters (must agree between builds)."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import cammiq_b200 as cq  # noqa: E402
from cammiq_b200 import synthlib as sl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--filter-mb", type=float, default=-1)
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--mode", default="p")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--random-reads", action="store_true",
                    help="uniform random reads unrelated to the genomes: (almost) no candidates, so the kernel's L2 traffic is "
                         "the filter probes plus the tile words -- isolates the filter's L2 hit rate under ncu")
    ap.add_argument("--workdir", default=os.environ.get("CAMMIQ_BENCH_DIR", "/tmp"))
    a = ap.parse_args()
    w = dict(bench.WORKLOADS[a.workload])
    if a.reads:
        w["reads"] = a.reads
    d = bench.ensure_index(a.workload, w, a.workdir)
    idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
    if a.filter_mb >= 0:
        idx.set_filter_budget(int(a.filter_mb * (1 << 20)))
    ctx = cq.Context(0).upload(idx, w["n_genomes"])
    n, rl = w["reads"], w["read_len"]
    if a.random_reads:
        rng = np.random.default_rng(12345)
        reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, rl), dtype=np.uint8)]
    else:
        reads = sl.make_reads(bench.synth_params(w), 0, n, rl, w["erate"])
    lengths = np.full(n, rl, np.uint8)
    if a.packed:
        pk, pl, _ = cq.pack_reads(reads.reshape(-1), None, lengths, stride=rl, threads=8)
        ctx.stage_packed(pk.reshape(-1), None, pl, stride=pk.shape[1])
    else:
        ctx.stage(reads.reshape(-1), None, lengths, stride=rl)
    mode = cq.MODE_P if a.mode == "p" else cq.MODE_SC
    ms, pk = [], []
    warm = 0 if a.iters <= 1 else 40   # ~0.3 s of launches: clocks and power state settle before timing
    for i in range(a.iters + warm):
        ctx.reset()
        if i >= warm:
            ctx.timing_reset()
        ctx.query_staged(mode)
        if i >= warm:
            ctx.sync()
            t = ctx.timing()
            ms.append(t["scan_ms_sum"])
            pk.append(t["pack_ms_sum"])
    r = ctx.fetch(mode)
    chk = [int(r["nundet"]), int(r["nconf"]), int(r["cnt_u"].sum()), int(r["cnt_d"].sum())]
    if mode == cq.MODE_P:
        chk += [int(r["rcount_u"].sum()), int(r["rcount_d"].sum()),
                int((r["rcount_u"].astype(np.uint64) * (np.arange(len(r["rcount_u"]), dtype=np.uint64) % 1000003)).sum())]
    print(json.dumps({"lib": os.path.basename(cq.capi.library_path()), "workload": a.workload, "reads": n, "read_len": rl,
                      "packed": a.packed, "random_reads": a.random_reads, "filter_mb": idx.info.filter_bytes / (1 << 20),
                      "scan_ms_mean": float(np.mean(ms)), "scan_ms_min": float(np.min(ms)), "pack_ms_mean": float(np.mean(pk)),
                      "reads_per_s": n / (float(np.mean(ms)) * 1e-3),
                      "regs": t["regs_per_thread"], "grid": t["grid_blocks"], "blocks_per_sm": t["blocks_per_sm"],
                      "dyn_smem": t["dyn_smem_bytes"], "probes": t["probes"], "candidates": t["bucket_hits"],
                      "leaf_hits": t["leaf_hits"], "chained": t["chained_loads"], "checksum": chk}))


if __name__ == "__main__":
    main()
