import sys, json
sys.path.insert(0, "/root/repo")
import cammiq_b200 as cq
ctx = cq.Context(0)
rows = []
for gb in (1, 4, 16, 32, 64):
    g = ctx.bench_random_gather(gb << 30, 32, 1 << 28, iters=2)
    rows.append({"region_gb": gb, "access_bytes": 32, "gaccess_per_s": g})
    print(rows[-1], flush=True)
json.dump(rows, open("/root/repo/gpurun_out/microbench_gather_large.json", "w"), indent=1)
