"""Index load / GPU context / upload wall times on the GPU box (cfg2 index); CAMMIQ_VERBOSE=1 adds
the per-phase lines of the decoder, the table build and the upload."""
import sys, time, os
sys.path.insert(0, "/root/repo")
import cammiq_b200 as cq
from cammiq_b200 import synthlib as sl
d = "/tmp/cli_e2e"
p = sl.params(seed=2, n_genomes=500, genome_len=3_000_000, cluster_size=4)
if not os.path.exists(d + "/index_u.bin1"):
    sl.write_index(p, d)
t = time.time(); idx = cq.Index(d + "/index_u.bin1", d + "/index_d.bin2"); print("load %.3f" % (time.time() - t), flush=True)
t = time.time(); ctx = cq.Context(0); print("ctx %.3f" % (time.time() - t), flush=True)
for i in range(3):
    t = time.time(); ctx.upload(idx, 500); print("upload %.3f" % (time.time() - t), flush=True)
