"""Wall-clock comparison of the drop-in CLI and the reference CLI on the same files (GPU box):
cfg2 index, FASTQ of N reads.  Writes gpurun_out/cli_e2e.json."""
import json
import os
import re
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cammiq_b200 import synthlib as sl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = "/tmp/cli_e2e"
p = sl.params(seed=2, n_genomes=500, genome_len=3_000_000, cluster_size=4)
if not os.path.exists(d + "/index_u.bin1"):
    sl.write_index(p, d)
fq = d + "/sample_%d.fq" % n
if not os.path.exists(fq):
    sl.write_fastq(p, 0, n, 100, 0.01, fq)
out = {"reads": n, "fastq_bytes": os.path.getsize(fq)}
arms = [("gpu_cli", os.path.join(REPO, "cammiq_b200", "cammiq"), [])]
if not os.environ.get("CLI_E2E_SKIP_REF"):
    arms.append(("reference_cli", os.path.join(REPO, "oracle", "_ref", "cammiq_ref"), ["-t", str(os.cpu_count())]))
for name, exe, extra in arms:
    cmd = [exe, "--query", "--read_cnts", "-f", d + "/genome_map.out", "-q", fq, "-i", d + "/index_u.bin1",
           d + "/index_d.bin2", "-o", d + "/" + name + ".out"] + extra
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, CAMMIQ_VERBOSE="1"))
    wall = time.time() - t
    err = r.stderr.replace("\r", "\n")
    g = lambda pat: (re.search(pat, err) or [None, None])[1]
    out[name] = {"wall_s": wall, "rc": r.returncode, "load_index_ms": g(r"Time for loading index: (\d+) ms"),
                 "query_ms": g(r"Time for query: (\d+) ms"), "nundet": g(r"unlabeled reads: (\d+)"),
                 "nconf": g(r"conflict labels: (\d+)"),
                 "verbose": [l for l in err.split("\n") if l.startswith("[")]}
# FASTQ ingest alone (parallel parse + 2-bit packing), from the CLI's own diagnostic
r = subprocess.run([os.path.join(REPO, "cammiq_b200", "cammiq"), "--dump_reads", fq], stdout=subprocess.DEVNULL,
                   stderr=subprocess.PIPE, text=True)
out["fastq_ingest"] = r.stderr.strip()
if len(arms) > 1:
    out["outputs_identical"] = open(d + "/gpu_cli.out").read() == open(d + "/reference_cli.out").read()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cli_e2e.json", "w"), indent=1)
print(json.dumps(out, indent=1))
