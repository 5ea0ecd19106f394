"""Where a cq_query call spends its wall time (cfg2 reads, both modes): host packing, the wait for
the scans, the fetch.  Runs on the GPU box; prints one line per mode."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import cammiq_b200 as cq  # noqa: E402
from cammiq_b200 import synthlib as sl  # noqa: E402

d = "/tmp/cli_e2e"
p = sl.params(seed=2, n_genomes=500, genome_len=3_000_000, cluster_size=4)
if not os.path.exists(d + "/index_u.bin1"):
    sl.write_index(p, d)
idx = cq.Index(d + "/index_u.bin1", d + "/index_d.bin2")
ctx = cq.Context(0).upload(idx, 500)
n, rl = 10_000_000, 100
host = torch.empty((n, rl), dtype=torch.uint8, pin_memory=True)
reads = host.numpy()
sl.make_reads(p, 0, n, rl, 0.01, out=reads)
lengths = torch.full((n,), rl, dtype=torch.uint8).pin_memory().numpy()
bufs = ctx.pinned_result_buffers()
for mode, name in ((cq.MODE_P, "P"), (cq.MODE_SC, "SC")):
    for rep in range(4):
        ctx.reset()
        t = time.perf_counter()
        ctx.query(mode, reads.reshape(-1), None, lengths, stride=rl, buffers=bufs)
        wall = (time.perf_counter() - t) * 1e3
        tm = ctx.timing()
    print("%s: wall %.2f ms | library total %.2f, host pack %.2f, fetch (d2h) %.2f, scan bracket %.2f | python %.2f" % (
        name, wall, tm["total_ms"], tm["host_pack_ms"], tm["d2h_ms"], tm["scan_ms"], wall - tm["total_ms"]))
