#!/usr/bin/env python
"""Random 8-byte gathers over a 64 MB (L2-resident) region with part of each SM's unified
L1 / shared memory reserved: does the gather rate depend on the L1 left over?  One process per
setting (the carve-out is fixed per kernel once set)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import cammiq_b200 as cq
ctx = cq.Context(0)
print(json.dumps({"g_per_s": ctx.bench_random_gather(64 << 20, 8, 1 << 28, iters=3)}))
''' % REPO

rows = []
for blocks, smem in ((8, 0), (8, 8 << 10), (8, 16 << 10), (8, 24 << 10), (8, 27 << 10), (3, 0), (3, 40 << 10), (3, 59 << 10),
                     (3, 72 << 10), (2, 0), (2, 59 << 10), (2, 100 << 10)):
    env = dict(os.environ, CAMMIQ_GATHER_SMEM=str(smem), CAMMIQ_GATHER_BLOCKS=str(blocks))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    try:
        v = json.loads(r.stdout.strip().splitlines()[-1])["g_per_s"]
    except Exception:
        v = None
    rows.append({"blocks_per_sm": blocks, "smem_per_block": smem, "smem_per_sm_kb": blocks * smem / 1024, "g_gathers_per_s": v})
    print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "microbench_l1.json"), "w"), indent=1)
