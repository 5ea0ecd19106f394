"""Random-access microbenchmark on the GPU box: gather rate (G accesses/s) versus region size
and access width -- the measured lookup roofline and the L2 capacity
curve that sizes the resident filter.  Writes gpurun_out/microbench_gather.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cammiq_b200 as cq  # noqa: E402

ctx = cq.Context(0)
out = {"device": ctx.device_info(), "rows": []}
print(out["device"], flush=True)
N = 1 << 28
for mb in (8, 16, 32, 48 + 16, 128, 256, 1024, 4096):
    size = mb << 20
    size = 1 << (size.bit_length() - 1)
    for acc in (4, 8, 32):
        for persist in ((False, True) if 32 <= mb <= 128 else (False,)):
            g = ctx.bench_random_gather(size, acc, N, iters=2, persist=persist)
            row = {"region_mb": size >> 20, "access_bytes": acc, "persist": persist, "gaccess_per_s": g,
                   "gb_per_s": g * acc}
            out["rows"].append(row)
            print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench_gather.json", "w"), indent=1)
