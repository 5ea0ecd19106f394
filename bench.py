#!/usr/bin/env python
"""bench.py -- reads/s of the CAMMiQ read-matching hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU scan

A step = one pass of the hot path (pack -> scan -> count reduction [-> NCCL reduce]) over one
batch of synthetic reads.  Headline workload = BASELINE.json configs[1]: `--both` index over 500
synthetic strain genomes (~1.5 Gbp), 10M simulated 100-bp reads with 1% substitutions.  N>1: index
replicated per GPU, every rank scans its own 10M reads (weak scaling), counters combined with one
grouped NCCL reduce per step.  ONE JSON line on stdout (rank 0).

Besides the headline the line carries, under "secondary", one block per other named shape of
BASELINE.json, each with its own throughput and a full-vector parity verdict:
    N=1:  cfg1_refbuilt  configs[0] on an index built by the reference's own builder (every counter,
                         every per-leaf rcount and the pair map against the reference's dump)
          cfg4           configs[3] shape: 5000 genomes, index too large for the L2 filter
          cfg5, cfg5_deep  configs[4]: 250-bp reads with N, 7% errors, near-duplicate strains; the
                         second index has every key longer than the hash length (h < k)
    N>1:  cfg3_sc_strong configs[2]: query64_sc path, 150-bp reads, 50M reads sharded over the ranks
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(n_genomes=500, genome_len=3_000_000, cluster_size=4, reads=10_000_000, read_len=100,
                 erate=0.01, k=26, lmax=50, seed=2, sample_reads=200_000,
                 desc="--both index, 500 synthetic strain genomes x 3 Mbp (1.5 Gbp), 10M simulated "
                      "100bp reads, 1% substitutions"),
    # configs[0] shape on the synthetic index writer (smoke size); the reference-built variant is
    # the cfg1_refbuilt block
    "cfg1": dict(n_genomes=10, genome_len=1_000_000, cluster_size=3, reads=100_000, read_len=100,
                 erate=0.0, k=26, lmax=50, seed=1, sample_reads=100_000,
                 desc="10 synthetic 1 Mbp genomes, 100k error-free 100bp reads"),
    # configs[2]: 150 bp reads on the cfg2 index
    "cfg3": dict(n_genomes=500, genome_len=3_000_000, cluster_size=4, reads=10_000_000, read_len=150,
                 erate=0.01, k=26, lmax=50, seed=2, sample_reads=200_000,
                 desc="cfg2 index, 150bp reads"),
    # configs[3] shape: 5000 genomes, ~15 Gbp
    "cfg4": dict(n_genomes=5000, genome_len=3_000_000, cluster_size=4, reads=4_000_000, read_len=150,
                 erate=0.01, k=26, lmax=50, seed=4, sample_reads=20_000,
                 desc="5000 synthetic genomes x 3 Mbp (15 Gbp): index too large for the L2 filter, 150bp reads"),
    # configs[4]: near-duplicate strains (pair-shared sequence dominates: D-heavy), half the keys
    # longer than h, 250-bp reads with 7% substitutions and N
    "cfg5": dict(n_genomes=200, genome_len=2_000_000, cluster_size=2, reads=2_000_000, read_len=250,
                 erate=0.07, n_rate=0.01, k=26, lmax=50, seed=5, permille_deep=500, permille_private=60,
                 permille_pair=800, sample_reads=50_000,
                 desc="adversarial: 200 near-duplicate strain genomes (80% pair-shared), half the keys deeper than h, "
                      "250bp reads, 7% substitutions, 1% N (substituted on the host)"),
    # the same with h < k: every key is longer than the hash length (tries of 1..30 levels)
    "cfg5_deep": dict(n_genomes=200, genome_len=2_000_000, cluster_size=2, reads=2_000_000, read_len=250,
                      erate=0.07, n_rate=0.01, k=20, lmax=50, seed=6, permille_deep=1000, permille_private=60,
                      permille_pair=800, sample_reads=50_000,
                      desc="adversarial, h=20 < every key length (21..50): each hit walks the trie; 250bp, 7% subs, 1% N"),
    # hit-uniqueness stress: a dense index (600 keys per kbp block) and nearly error-free 250-bp reads:
    # about a hundred leaf hits per read, many of them the same leaf on both strands' positions
    "cfg5_dense": dict(n_genomes=64, genome_len=1_000_000, cluster_size=2, reads=500_000, read_len=250,
                       erate=0.005, k=26, lmax=50, seed=7, permille_deep=300, permille_private=150,
                       permille_pair=800, u_per_block=600, d_per_block=600, sample_reads=20_000,
                       desc="adversarial, hit lists: 64 near-duplicate strains, 600 keys per 1024-base block, 250bp reads, 0.5% subs"),
}
STRONG_CFG3_READS = 50_000_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_params(w, threads=0):
    from cammiq_b200 import synthlib as sl
    extra = {k: w[k] for k in ("permille_private", "permille_pair", "u_per_block", "d_per_block") if k in w}
    return sl.params(seed=w["seed"], n_genomes=w["n_genomes"], genome_len=w["genome_len"],
                     cluster_size=w["cluster_size"], k=w["k"], lmax=w["lmax"], threads=threads,
                     permille_deep=w.get("permille_deep", 50), **extra)


def workdir_for(name, w, base):
    tag = "_".join("%s%d" % (k[:2], w[k]) for k in ("permille_private", "permille_pair") if k in w)
    return os.path.join(base, "cammiq_bench_%s_g%d_l%d_s%d_d%d_k%d%s" % (
        name, w["n_genomes"], w["genome_len"], w["seed"], w.get("permille_deep", 50), w["k"], "_" + tag if tag else ""))


def ensure_index(name, w, base):
    """Synthetic index files (format-exact .bin1/.bin2 + map + meta), generated once."""
    from cammiq_b200 import synthlib as sl
    if name.startswith("cfg3"):
        name, w = "cfg2", dict(WORKLOADS["cfg2"])   # configs[2] runs on the configs[1] index
    d = workdir_for(name, w, base)
    done = os.path.join(d, ".done")
    if not os.path.exists(done):
        t = time.time()
        st = sl.write_index(synth_params(w), d)
        open(done, "w").write(json.dumps(st))
        log("[bench] synthetic index written to %s in %.1f s: %s" % (d, time.time() - t, st))
    return d


def make_reads(w, first, n, out=None):
    """Simulated reads of a workload; N (n_rate) is inserted and then substituted the way the
    reference's reader does it -- one random base per read (query.cpp:383) -- with a seeded
    generator, so that every consumer sees the same post-substitution reads."""
    from cammiq_b200 import synthlib as sl
    reads = sl.make_reads(synth_params(w), first, n, w["read_len"], w["erate"], out=out)
    if w.get("n_rate", 0) > 0:
        # per block of 65536 reads, seeded by the block number: the same read gets the same N
        # positions and the same substitute whatever range it is generated in
        rl, blk = w["read_len"], 65536
        alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
        for b in range(first // blk, (first + n + blk - 1) // blk):
            rng = np.random.default_rng([w["seed"], 7919, b])
            mask = rng.random((blk, rl)) < w["n_rate"]
            sub = alpha[rng.integers(0, 4, blk)]
            lo, hi = max(first, b * blk), min(first + n, (b + 1) * blk)
            m = mask[lo - b * blk:hi - b * blk]
            view = reads[lo - first:hi - first]
            view[m] = np.broadcast_to(sub[lo - b * blk:hi - b * blk, None], m.shape)[m]
    return reads


def write_sample_fastq(reads, path):
    """FASTQ of a read array (for the reference harness)."""
    n, rl = reads.shape
    qual = b"I" * rl
    with open(path + ".tmp", "wb") as f:
        for i in range(n):
            f.write(b"@r%d\n" % i)
            f.write(reads[i].tobytes())
            f.write(b"\n+\n")
            f.write(qual)
            f.write(b"\n")
    os.rename(path + ".tmp", path)


def ensure_sample_fastq(name, w, d):
    fq = os.path.join(d, "sample_%s_%d_%d.fq" % (name, w["sample_reads"], w["read_len"]))
    if not os.path.exists(fq):
        write_sample_fastq(make_reads(w, 0, w["sample_reads"]), fq)
    return fq


def bytes_per_read(rl, h, hits_per_read, rcount_updates_per_read):
    """Algorithmic bytes per read of the scan kernel, SURVEY.md section 8d:
    B = 32*P + 32*H + 2*32*A + ceil(rl/4) + 16, P = 2*(rl-h+1) merged-table probes."""
    P = 2 * (rl - h + 1)
    return 32.0 * P + 32.0 * hits_per_read + 64.0 * rcount_updates_per_read + (rl + 3) // 4 + 16


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Several ranks share a host: keep a rank's threads (the host packer) and its page-locked buffers on the
    NUMA node its GPU hangs off.  Best effort; returns a description for the config block."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf            # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "gpu %d: no NUMA node reported" % local
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        mine = cpus & os.sched_getaffinity(0)
        if len(mine) < 2:
            return "gpu %d on node %d: fewer than 2 of its cpus in this process's mask" % (local, node)
        os.sched_setaffinity(0, mine)
        return "gpu %d -> NUMA node %d, %d cpus" % (local, node, len(mine))
    except Exception as e:
        return "not bound (%s)" % str(e)[:80]


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------- the reference on the host cores

def run_reference_harness(d, fq, threads, reps, mode="mt"):
    """The UNMODIFIED reference scan (oracle/_ref/ref_harness) on the host cores; one dict per
    repetition with the timing, the full per-genome vectors and the per-leaf rcount digests."""
    exe = os.path.join(REPO, "oracle", "_ref", "ref_harness")
    if not os.access(exe, os.X_OK):
        return None
    cmd = [exe, "time", os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"),
           os.path.join(d, "genome_map.out"), mode, str(threads), fq, str(reps)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    if res.returncode != 0 or not rows:
        log("[bench] ref_harness failed: rc=%d %s" % (res.returncode, res.stderr[-400:]))
        return None
    return rows


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wraps modulo 2^64) -- oracle/ref_harness.cpp mix64."""
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def rcount_digest(idx, table, G, rcount):
    """(sum, digest) of the per-leaf counts over Hash::map_sp in its own order, as ref_harness
    computes them on the reference's nodes: sum over g, k of rcount(map_sp[g][k]) * mix64(g<<32 | k)."""
    off, ids = idx.map_sp(table, G)
    off = off.astype(np.int64)
    counts = np.diff(off[1:G + 2])
    g = np.repeat(np.arange(1, G + 1, dtype=np.uint64), counts)
    k = np.arange(len(ids), dtype=np.uint64) - np.repeat(off[1:G + 1], counts).astype(np.uint64) + np.uint64(off[1])
    rc = rcount[ids.astype(np.int64)].astype(np.uint64)
    with np.errstate(over="ignore"):
        w = mix64((g << np.uint64(32)) | k)
        return int(rc.sum()), "%016x" % int((rc * w).sum(dtype=np.uint64))


def compare_with_harness(idx, G, got, row, mode):
    """Full-vector parity of a GPU result against one ref_harness repetition."""
    bad = []
    if int(got["nundet"]) != row["nundet"] or int(got["nconf"]) != row["nconf"]:
        bad.append("nundet/nconf %d/%d vs %d/%d" % (got["nundet"], got["nconf"], row["nundet"], row["nconf"]))
    if [int(x) for x in got["cnt_u"][1:G + 1]] != row["cu"]:
        bad.append("cnt_u[]")
    if [int(x) for x in got["cnt_d"][1:G + 1]] != row["cd"]:
        bad.append("cnt_d[]")
    if mode == "sc":
        want = {(a, b): c for a, b, c in row.get("pairs", [])}
        if got["pairs"] != want:
            bad.append("pair map")
    else:
        import cammiq_b200 as cq
        for tag, table, key in (("rcu", cq.TABLE_U, "rcount_u"), ("rcd", cq.TABLE_D, "rcount_d")):
            s, dg = rcount_digest(idx, table, G, got[key])
            if s != row[tag + "_sum"] or dg != row[tag + "_digest"]:
                bad.append("%s sum %d digest %s vs %d %s" % (key, s, dg, row[tag + "_sum"], row[tag + "_digest"]))
    return "ok" if not bad else "MISMATCH: " + "; ".join(bad)


def impl_reference(args, w, name, json_fd):
    """Reference arm: the reference's own OpenMP scan (query64mt_p) on this box's host cores,
    on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    d = ensure_index(name, w, args.workdir)
    fq = ensure_sample_fastq(name, w, d)
    threads = host_threads()
    rows = run_reference_harness(d, fq, threads, args.warmup + args.steps)
    kind = "reference"
    if rows is None:
        rows, kind = oracle_port_rows(d, w, args.warmup + args.steps), "port"
        threads = 1
    timed = rows[args.warmup:]
    ms = float(np.mean([r["query_ms"] for r in timed]))
    value = w["sample_reads"] / (ms * 1e-3)
    sample = "%d of the workload's %d reads per step (first reads of the same seeded stream), full index" % (
        w["sample_reads"], w["reads"])
    line = {
        "impl": "reference", "metric": "reads/sec classified", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": name + ": " + w["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": kind, "sample": sample,
                         "note": "reference query64mt_p compiled with std::unordered_map standing in for "
                                 "robin_hood (not vendored); index load %.1f s excluded" % (rows[0].get("load_ms", 0) / 1e3)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json_fd, line)
    return 0


def oracle_port_rows(d, w, reps):
    """Fallback CPU arm when oracle/_ref is absent: the C restatement, one thread."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    n = min(w["sample_reads"], 50_000)
    reads = make_reads(w, 0, n)
    ou, od = ol.OracleIndex(os.path.join(d, "index_u.bin1")), ol.OracleIndex(os.path.join(d, "index_d.bin2"))
    lens = np.full(n, w["read_len"], np.uint8)
    offs = np.arange(n, dtype=np.uint64) * w["read_len"]
    rows = []
    for _ in range(reps):
        t = time.time()
        ol.oracle_query(ou, od, ol.MODE_P, w["n_genomes"], reads.reshape(-1), offs, lens)
        rows.append({"query_ms": (time.time() - t) * 1e3 * (w["sample_reads"] / n), "load_ms": 0})
    return rows


# ------------------------------------------------------------------------------- roofline pieces

def git_blob_hash(path):
    data = open(path, "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def profiled_counters(name):
    """ncu figures of the scan kernel (profiles/scan_traffic.json), valid only for the kernel source
    they were captured from: the file carries the git blob hash of scan_kernels.cuh."""
    try:
        t = json.load(open(os.path.join(REPO, "profiles", "scan_traffic.json")))
    except Exception:
        return None, "profiles/scan_traffic.json missing"
    have = git_blob_hash(os.path.join(REPO, "cammiq_b200", "csrc", "scan_kernels.cuh"))
    if t.get("scan_kernels_cuh_blob") != have:
        return None, "profiles/scan_traffic.json was captured from another scan_kernels.cuh (%s, now %s)" % (
            str(t.get("scan_kernels_cuh_blob"))[:12], have[:12])
    return t.get(name), None


def gather_peak(ctx, region_bytes, smem_per_block=0, blocks_per_sm=8, carveout_pct=None):
    """Random 8-byte gathers over an L2-sized region (what the filter probes are), optionally with
    the shared-memory footprint and carve-out of the scan kernel taken out of each SM's L1."""
    r = 1 << 20
    while r < region_bytes:
        r <<= 1
    os.environ["CAMMIQ_GATHER_SMEM"] = str(int(smem_per_block))
    os.environ["CAMMIQ_GATHER_BLOCKS"] = str(int(blocks_per_sm))
    if carveout_pct is not None:
        os.environ["CAMMIQ_GATHER_CARVEOUT"] = str(int(carveout_pct))
    try:
        return ctx.bench_random_gather(r, 8, 1 << 27, iters=2)
    finally:
        for k in ("CAMMIQ_GATHER_SMEM", "CAMMIQ_GATHER_BLOCKS", "CAMMIQ_GATHER_CARVEOUT"):
            os.environ.pop(k, None)


def timed_scan(ctx, mode, steps, warmup):
    """CUDA-event time of `steps` resident passes (reset + pack + scan + count reduction)."""
    for _ in range(warmup):
        ctx.reset()
        ctx.query_staged(mode)
    ctx.sync()
    ctx.timing_reset()
    l0 = ctx.timing()["kernel_launches"]
    for _ in range(steps):
        ctx.reset()
        ctx.query_staged(mode)
    ctx.sync()
    t = ctx.timing()
    n = max(t["steps"], 1)
    return {"scan_ms": t["scan_ms_sum"] / n, "pack_ms": t["pack_ms_sum"] / n, "launches": t["kernel_launches"] - l0, "timing": t}


# ------------------------------------------------------------------------------- secondary blocks

def parse_ref_dump(path):
    """oracle/ref_harness dump -> per-genome vectors, per-leaf rcount lists in map_sp order, pair map."""
    d = {"rcu": {}, "rcd": {}, "pairs": {}}
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "G":
            d["g"] = int(t[1])
        elif t[0] in ("NUNDET", "NCONF"):
            d[t[0].lower()] = int(t[1])
        elif t[0] in ("CU", "CD"):
            d[t[0].lower()] = [int(x) for x in t[1:]]
        elif t[0] in ("RCU", "RCD"):
            d[t[0].lower()][int(t[1])] = [int(x) for x in t[3:]]
        elif t[0] == "PAIRS":
            for item in t[2:]:
                a, b, c = item.split(":")
                d["pairs"][(int(a), int(b))] = int(c)
        elif t[0] == "READ":
            break
    return d


def read_fastq_array(path):
    seqs = [l.rstrip(b"\n") for i, l in enumerate(open(path, "rb")) if i % 4 == 1]
    lengths = np.array([len(s) for s in seqs], dtype=np.uint8)
    offsets = np.zeros(len(seqs), dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths[:-1].astype(np.uint64))
    return np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(), offsets, lengths


def block_cfg1_refbuilt(cq, steps, warmup):
    """configs[0] on a REFERENCE-BUILT index (oracle/fixtures/make_cfg1.py: cammiq --build --both by the
    unmodified builder, 10 x 1 Mbp, 100k error-free reads): every counter, every per-leaf rcount in
    map_sp order and the pair map against the reference's own dumps."""
    d = os.path.join(REPO, "oracle", "_ref", "fixtures", "cfg1")
    if not os.path.exists(os.path.join(d, ".done")):
        mk = os.path.join(REPO, "oracle", "fixtures", "make_cfg1.py")
        if os.access(os.path.join(REPO, "oracle", "_ref", "cammiq_ref"), os.X_OK):
            subprocess.run([sys.executable, mk], capture_output=True)
    if not os.path.exists(os.path.join(d, ".done")):
        return {"skipped": "oracle/_ref/fixtures/cfg1 absent (built by oracle/fixtures/make_cfg1.py where /root/reference exists)"}
    idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
    dump = {m: parse_ref_dump(os.path.join(d, "dump_%s.txt" % m)) for m in ("p", "sc")}
    G = dump["p"]["g"]
    ctx = cq.Context(0).upload(idx, G)
    bases, offsets, lengths = read_fastq_array(os.path.join(d, "reads.fq"))
    out = {"workload": "BASELINE configs[0]: reference-built --both index over 10 x 1 Mbp strain genomes, "
                       "%d error-free 100bp reads" % len(lengths),
           "leaves_u": idx.n_leaves_u, "leaves_d": idx.n_leaves_d, "index_built_by": "oracle/_ref/cammiq_ref --build --both"}
    verdicts = []
    for mode, m in ((cq.MODE_P, "p"), (cq.MODE_SC, "sc")):
        got = ctx.query(mode, bases, offsets, lengths)
        ctx.reset()
        ref = dump[m]
        bad = []
        if (int(got["nundet"]), int(got["nconf"])) != (ref["nundet"], ref["nconf"]):
            bad.append("nundet/nconf")
        if [int(x) for x in got["cnt_u"][1:]] != ref["cu"] or [int(x) for x in got["cnt_d"][1:]] != ref["cd"]:
            bad.append("cnt vectors")
        if m == "p":
            n_checked = 0
            for tag, table, key in (("rcu", cq.TABLE_U, "rcount_u"), ("rcd", cq.TABLE_D, "rcount_d")):
                off, ids = idx.map_sp(table, G)
                for g in range(1, G + 1):
                    mine = [int(x) for x in got[key][ids[int(off[g]):int(off[g + 1])].astype(np.int64)]]
                    if mine != ref[tag].get(g, []):
                        bad.append("%s of genome %d" % (key, g))
                    n_checked += len(mine)
            out["rcount_entries_compared"] = n_checked
        elif got["pairs"] != ref["pairs"]:
            bad.append("pair map")
        verdicts.append("query64_%s: %s" % (m, "ok" if not bad else "MISMATCH " + ", ".join(bad[:4])))
    ctx.stage(bases, offsets, lengths)
    t = timed_scan(ctx, cq.MODE_P, steps, warmup)
    out["reads_per_s"] = len(lengths) / ((t["scan_ms"] + t["pack_ms"]) * 1e-3)
    out["ms_per_step"] = t["scan_ms"] + t["pack_ms"]
    out["parity"] = "ok" if all(v.endswith("ok") for v in verdicts) else "; ".join(verdicts)
    out["parity_detail"] = verdicts
    out["compared"] = "nundet, nconf, cnt_u[], cnt_d[], every per-leaf rcount in map_sp order, pair map -- reference dumps"
    ctx.close()
    return out


def oracle_sample_check(cq, ctx, idx_dir, G, reads, lengths, rl, m, mode_pair):
    """Per-read parity on the first m reads against the C restatement (oracle/): class, genome ids,
    counters, per-leaf rcount."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    t = time.time()
    ou, od = ol.OracleIndex(os.path.join(idx_dir, "index_u.bin1")), ol.OracleIndex(os.path.join(idx_dir, "index_d.bin2"))
    load_s = time.time() - t
    offs = np.arange(m, dtype=np.uint64) * rl
    verdict = []
    t = time.time()
    for cmode, omode, tag in ((cq.MODE_P, ol.MODE_P, "p"), (cq.MODE_SC, ol.MODE_SC, "sc")):
        if tag not in mode_pair:
            continue
        o = ol.oracle_query(ou, od, omode, G, reads[:m].reshape(-1), offs, lengths[:m], per_read=True)
        ctx.reset()
        g = ctx.query(cmode, reads[:m].reshape(-1), None, lengths[:m], stride=rl, per_read=True)
        ctx.reset()
        keys = ["cnt_u", "cnt_d", "read_class", "read_rid_a", "read_rid_b"] + (["rcount_u", "rcount_d"] if tag == "p" else [])
        ok = all(np.array_equal(np.asarray(o[k]), np.asarray(g[k])) for k in keys)
        ok = ok and o["nundet"] == g["nundet"] and o["nconf"] == g["nconf"] and (tag == "p" or o["pairs"] == g["pairs"])
        verdict.append("%s: %s" % (tag, "ok" if ok else "MISMATCH"))
    return {"reads": m, "verdict": verdict, "oracle_load_s": load_s, "oracle_query_s": time.time() - t,
            "ok": all(v.endswith("ok") for v in verdict)}


def block_synthetic(cq, name, workdir, steps, warmup, budget_left, modes=("p",), with_harness=True, dedup_ab=False):
    """A secondary shape on the synthetic index writer: throughput of the resident scan, the
    size-independent properties on all reads, per-read parity with the oracle and full-vector parity
    with the reference harness on a sample."""
    w = dict(WORKLOADS[name])
    t0 = time.time()
    d = ensure_index(name, w, workdir)
    idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
    info = idx.info
    G, n, rl = w["n_genomes"], w["reads"], w["read_len"]
    ctx = cq.Context(0).upload(idx, G)
    prep_s = time.time() - t0
    reads = make_reads(w, 0, n)
    lengths = np.full(n, rl, dtype=np.uint8)
    out = {"workload": name + ": " + w["desc"], "reads": n, "read_len": rl, "hash_len": info.hash_len,
           "leaves_u": info.n_leaves_u, "leaves_d": info.n_leaves_d, "trie_nodes": info.n_nodes_u + info.n_nodes_d,
           "table_gb": info.n_table_buckets * 32 / 1e9, "filter_mb": info.filter_bytes / (1 << 20),
           "index_device_gb": info.device_bytes / 1e9, "index_prepare_s": prep_s}
    ctx.stage(reads.reshape(-1), None, lengths, stride=rl)
    for tag in modes:
        mode = cq.MODE_P if tag == "p" else cq.MODE_SC
        t = timed_scan(ctx, mode, steps, warmup)
        tm = t["timing"]
        ms = t["scan_ms"] + t["pack_ms"]
        key = "" if tag == "p" else "_sc"
        out["reads_per_s" + key] = n / (ms * 1e-3)
        out["ms_per_step" + key] = ms
        out["scan_ms" + key], out["pack_ms" + key] = t["scan_ms"], t["pack_ms"]
        if tag == modes[0]:
            out["candidates_per_read"] = tm["bucket_hits"] / n
            out["leaf_hits_per_read"] = tm["leaf_hits"] / n
            out["chained_bucket_loads"] = tm["chained_loads"]
            out["launch"] = {"grid": tm["grid_blocks"], "blocks_per_sm": tm["blocks_per_sm"], "dyn_smem": tm["dyn_smem_bytes"],
                             "regs": tm["regs_per_thread"]}
            positions = tm["probes"] / 2
            sec = t["scan_ms"] * 1e-3
            if info.filter_bytes == 0 or tm["sieve_loads"] > 0:
                # the regime bound by random HBM accesses: one table sector per read position, or --
                # behind the sieve -- per position that passes it
                gsec = ctx.bench_random_sectors(1 << 28, iters=2)
                first = tm["sieve_loads"] if tm["sieve_loads"] > 0 else positions
                sectors = first + tm["chained_loads"] + tm["bucket_hits"] + tm["leaf_hits"]
                out["roofline"] = {"bound": "hbm-random", "unit": "G sectors/s", "peak": gsec,
                                   "peak_source": "cq_bench_random_sectors over this table, same run",
                                   "achieved": sectors / sec / 1e9, "frac": sectors / sec / 1e9 / gsec,
                                   "read_positions_g_per_s": positions / sec / 1e9,
                                   "sieve_pass_rate": (tm["sieve_loads"] / positions) if tm["sieve_loads"] > 0 else None,
                                   "sectors": {"phase1_bucket_keys": first, "chained_buckets": tm["chained_loads"],
                                               "phase2_buckets": tm["bucket_hits"], "leaf_records": tm["leaf_hits"]},
                                   "note": "achieved = random 32-byte sectors the scan requests from HBM (bucket keys of phase 1, "
                                           "chained buckets, phase-2 buckets, leaf records; rcount updates not counted) per second "
                                           "of scan kernel time, against the measured random-sector rate of this table"}
                if tm["sieve_loads"] > 0:
                    g = gather_peak(ctx, info.filter_bytes, carveout_pct=tm["smem_carveout_pct"],
                                    smem_per_block=tm["smem_carveout_pct"] * 228 * 1024 // 100 // max(tm["blocks_per_sm"], 1) - 1024,
                                    blocks_per_sm=tm["blocks_per_sm"])
                    out["roofline"]["l2_gather"] = {"sieve_loads_g_per_s": positions / sec / 1e9, "peak_g_per_s": g,
                                                    "frac": positions / sec / 1e9 / g}
    if dedup_ab:
        # the same scan with every hit list deduplicated by its own lane (the quadratic loop that
        # was the only path before the warp-cooperative hash set existed)
        os.environ["CAMMIQ_LIGHT_HITS"] = "100000"
        try:
            tq = timed_scan(ctx, cq.MODE_P, steps, warmup)
        finally:
            os.environ.pop("CAMMIQ_LIGHT_HITS", None)
        out["dedup"] = {"warp_cooperative_hash_set_scan_ms": out["scan_ms"], "per_lane_quadratic_scan_ms": tq["scan_ms"],
                        "speedup": tq["scan_ms"] / out["scan_ms"]}
    # size-independent properties on ALL reads: class counts add up, strand symmetry of the totals
    ctx.reset()
    a = ctx.query(cq.MODE_P, reads.reshape(-1), None, lengths, stride=rl, per_read=True)
    ctx.reset()
    cls = a["read_class"]
    cons = (int((cls == 0).sum()) == a["nundet"] and int((cls == 1).sum()) == a["nconf"]
            and int(a["cnt_u"].sum()) == int(((cls == 2) | (cls == 4)).sum())
            and int(a["cnt_d"].sum()) == int((cls == 3).sum()) * 2 + int(((cls == 4) | (cls == 5)).sum()))
    lut = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        lut[x] = y
    half = min(n, 500_000)
    rc = np.ascontiguousarray(lut[reads[:half, ::-1]])
    f = ctx.query(cq.MODE_P, reads[:half].reshape(-1), None, lengths[:half], stride=rl)
    ctx.reset()
    b = ctx.query(cq.MODE_P, rc.reshape(-1), None, lengths[:half], stride=rl)
    ctx.reset()
    sym = all(np.array_equal(f[k], b[k]) for k in ("cnt_u", "cnt_d", "rcount_u", "rcount_d")) and f["nundet"] == b["nundet"]
    out["properties"] = {"class_counts_add_up": bool(cons), "reverse_complemented_reads_same_totals": bool(sym),
                         "class_histogram": {int(k): int(v) for k, v in zip(*np.unique(cls, return_counts=True))}}
    verdicts = [cons, sym]
    m = min(w["sample_reads"], n, 20_000)
    if budget_left() > 60:
        oc = oracle_sample_check(cq, ctx, d, G, reads, lengths, rl, m, modes)
        out["oracle_sample"] = oc
        verdicts.append(oc["ok"])
    else:
        out["oracle_sample"] = {"skipped": "time budget"}
    if with_harness and budget_left() > 45:
        fq = ensure_sample_fastq(name, w, d)
        s = w["sample_reads"]
        hv = []
        for tag in modes:
            rows = run_reference_harness(d, fq, host_threads(), 1, mode="mt" if tag == "p" else "sc")
            if rows:
                mode = cq.MODE_P if tag == "p" else cq.MODE_SC
                g = ctx.query(mode, reads[:s].reshape(-1), None, lengths[:s], stride=rl)
                ctx.reset()
                v = compare_with_harness(idx, G, g, rows[-1], tag)
                hv.append("%s: %s" % (tag, v))
                verdicts.append(v == "ok")
                out.setdefault("cpu_reference_reads_per_s", {})[tag] = s / (rows[-1]["query_ms"] * 1e-3)
        out["reference_sample"] = {"reads": s, "verdict": hv, "compared": "nundet, nconf, cnt_u[], cnt_d[], rcount sums + "
                                   "position-weighted digests over map_sp (p) / pair map (sc)"}
    out["parity"] = "ok" if all(verdicts) else "MISMATCH"
    ctx.close()
    return out


def abi_multi_check(name, w, workdir, world, single):
    """cq_multi_* (reads sharded over the GPUs, grouped NCCL reduce -- all behind the C ABI) on a
    sample, in a process of its own with a timeout; must reproduce this rank's single-GPU totals."""
    code = r'''
import json, os, sys
sys.path.insert(0, %(repo)r)
import numpy as np
import bench
import cammiq_b200 as cq
w = json.loads(%(w)r)
d = bench.workdir_for(%(name)r, w, %(workdir)r)
idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
n, rl = %(n)d, w["read_len"]
reads = bench.make_reads(w, 0, n)
lengths = np.full(n, rl, np.uint8)
m = cq.MultiContext(%(world)d).upload(idx, w["n_genomes"])
out = {}
for mode, tag in ((cq.MODE_P, "p"), (cq.MODE_SC, "sc")):
    r = m.query(mode, reads.reshape(-1), None, lengths, stride=rl)
    m.reset()
    out[tag] = [int(r["nundet"]), int(r["nconf"]), [int(x) for x in r["cnt_u"]], [int(x) for x in r["cnt_d"]]]
    if tag == "p":
        out[tag] += [bench.rcount_digest(idx, cq.TABLE_U, w["n_genomes"], r["rcount_u"]), bench.rcount_digest(idx, cq.TABLE_D, w["n_genomes"], r["rcount_d"])]
    else:
        out[tag] += [sorted([a, b, c] for (a, b), c in r["pairs"].items())]
out["info"] = m.info()
print("RESULT " + json.dumps(out))
''' % dict(repo=REPO, w=json.dumps(w), name=name, workdir=workdir, n=single["n"], world=world)
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "LOCAL_WORLD_SIZE", "CUDA_VISIBLE_DEVICES"):
        if k != "CUDA_VISIBLE_DEVICES":
            env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240, env=env)
    except subprocess.TimeoutExpired:
        return "timeout"
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    if r.returncode != 0 or not line:
        return "failed: " + r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else "failed"
    got = json.loads(line[0][7:])
    info = got.pop("info")
    want = json.loads(json.dumps(single["result"]))
    if got != want:
        return "MISMATCH"
    return "ok (%d reads over %d GPUs %s, NCCL %s, reduce %.2f ms)" % (single["n"], info["n_gpus"], info["shard_reads"],
                                                                  info["nccl_version"], info["reduce_ms"])


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version
    # banner, ...) is diverted to stderr at the file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        return run(json_fd)
    finally:
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        os.close(json_fd)


def emit(json_fd, line):
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def run(json_fd):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cammiq", choices=["cammiq", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU per step")
    ap.add_argument("--workdir", default=os.environ.get("CAMMIQ_BENCH_DIR", "/tmp"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline workload only")
    ap.add_argument("--secondary", default="", help="comma-separated secondary blocks (default: all that apply)")
    ap.add_argument("--time-budget", type=float, default=420.0, help="seconds after which remaining secondary blocks are skipped")
    ap.add_argument("--mode", default="p", choices=["p", "sc"])
    ap.add_argument("--deep-permille", type=int, default=-1, help="override the share of keys longer than h (synthetic index)")
    ap.add_argument("--pack-threads", type=int, default=-1,
                    help="host threads packing reads to 2 bits before the copy in the e2e leg (0 = ASCII over PCIe, "
                         "-1 = min(16, host cpus / ranks))")
    ap.add_argument("--filter-mb", type=float, default=-1, help="override the membership-filter budget (MB, 0 = none)")
    args = ap.parse_args()
    t_start = time.time()
    args.warmup = max(args.warmup, 3) if args.impl == "cammiq" else max(args.warmup, 1)
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.deep_permille >= 0:
        w["permille_deep"] = args.deep_permille
    if args.reads:
        w["reads"] = args.reads
        w["sample_reads"] = min(w["sample_reads"], args.reads)
    if args.impl == "reference":
        return impl_reference(args, w, name, json_fd)

    import torch
    import torch.distributed as dist

    import cammiq_b200 as cq
    from cammiq_b200 import multigpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback.")
    torch.cuda.set_device(local)
    host_group = None
    ncpu_unbound = host_threads()
    numa = bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("CAMMIQ_NO_NUMA_BIND") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")   # host-side waits that must not occupy the GPUs
    mode = cq.MODE_P if args.mode == "p" else cq.MODE_SC

    # ---- workload: index (rank 0 writes, everyone loads), reads (each rank its own shard) ----
    t0 = time.time()
    if rank == 0:
        d = ensure_index(name, w, args.workdir)
    if world > 1:
        dist.barrier()
    d = workdir_for("cfg2" if name.startswith("cfg3") else name, WORKLOADS["cfg2"] if name.startswith("cfg3") else w, args.workdir)
    idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
    if args.filter_mb >= 0:
        idx.set_filter_budget(int(args.filter_mb * (1 << 20)))
    info = idx.info
    # everything (kernels, memsets, NCCL hand-off, the timing events) runs on one torch stream
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream().cuda_stream
    ctx = cq.Context(local, stream=stream).upload(idx, w["n_genomes"])
    t_index = time.time() - t0
    n, rl, h, G = w["reads"], w["read_len"], info.hash_len, w["n_genomes"]
    host = torch.empty((n, rl), dtype=torch.uint8, pin_memory=True)
    reads = host.numpy()
    make_reads(w, rank * n, n, out=reads)
    lengths_t = torch.full((n,), rl, dtype=torch.uint8).pin_memory()
    lengths = lengths_t.numpy()
    log("[bench] rank %d: index %d U + %d D leaves, table %.2f GB, %d reads ready (%.1f s)" % (
        rank, info.n_leaves_u, info.n_leaves_d, info.n_table_buckets * 32 / 1e9, n, time.time() - t0))

    def as_tensors(arrays):
        return tuple(torch.as_tensor(a, device="cuda") for a in arrays)

    # the context's two accumulator sets (cq_swap_accumulators): with several ranks the NCCL reduce of one
    # step's counters runs on a side stream while the next step packs and scans into the other set
    overlap = world > 1 and mode == cq.MODE_P and not os.environ.get("CAMMIQ_NO_REDUCE_OVERLAP")
    acc_sets = [as_tensors(ctx.device_counter_arrays())]
    if overlap:
        ctx.swap_accumulators()
        acc_sets.append(as_tensors(ctx.device_counter_arrays()))
        ctx.swap_accumulators()
    cur = [0]                                    # index of the set the context accumulates into
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream() if overlap else None
    ev_scanned = [torch.cuda.Event(), torch.cuda.Event()]
    ev_reduced = [None, None]

    def combine():
        # the one collective of the path: sum the counter block (and the per-leaf rcount in
        # mode P) into rank 0 over NCCL, one grouped launch
        counts, rc_u, rc_d = acc_sets[cur[0]]
        if mode == cq.MODE_P:
            multigpu.combine_counters(counts, rc_u, rc_d)
        else:
            multigpu.combine_counters(counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident throughput: reads already in HBM ---------------------------------
    ctx.stage(reads.reshape(-1), None, lengths, stride=rl)

    def step_resident():
        if overlap and mode == cq.MODE_P:
            # step k: [wait until the reduce that last read this set is done] reset, pack, scan on the
            # context's stream; then the set is handed to the side stream, which reduces it while step
            # k+1 runs on the other set.  Every step is still reduced in full.
            k = cur[0]
            if ev_reduced[k] is not None:
                main_stream.wait_event(ev_reduced[k])
            ctx.reset()
            ctx.query_staged(mode)
            ev_scanned[k].record(main_stream)
            ctx.swap_accumulators()
            with torch.cuda.stream(side_stream):
                side_stream.wait_event(ev_scanned[k])
                multigpu.combine_counters(*acc_sets[k])
                if ev_reduced[k] is None:
                    ev_reduced[k] = torch.cuda.Event()
                ev_reduced[k].record(side_stream)
            cur[0] = k ^ 1
            return
        ctx.reset()
        ctx.query_staged(mode)
        combine()

    def join_reduces():
        # the timed region ends when the last step's reduce has: the context's stream waits for the side stream
        if overlap:
            main_stream.wait_stream(side_stream)

    # nvidia-smi needs a few hundred ms to come up and samples every 100 ms; the timed region is a few
    # tens of ms.  The sampler therefore runs while the SAME steps keep the GPU under load: untimed steps
    # until its first sample has arrived, the timed steps, untimed steps until three samples are in.
    sampler = ClockSampler(local)
    sampler.start()

    def keep_loaded(want_samples, max_s):
        t_start = time.time()
        while True:
            for _ in range(20):
                step_resident()
            join_reduces()
            torch.cuda.synchronize()
            more = rank == 0 and len(sampler.rows) < want_samples and time.time() - t_start < max_s
            if world > 1:   # rank 0 decides for everyone: the steps contain collectives
                f = torch.tensor([1 if more else 0], device="cuda")
                dist.all_reduce(f, op=dist.ReduceOp.MAX)
                more = bool(f.item())
            if not more:
                return

    for _ in range(args.warmup):
        step_resident()
    keep_loaded(1, 4.0)
    barrier()
    ctx.timing_reset()
    launches0 = ctx.timing()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    join_reduces()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    tm = ctx.timing()
    launches = tm["kernel_launches"] - launches0
    keep_loaded(3, 2.0)
    clocks = sampler.stop()
    clocks["note"] = "sampled every 100 ms while untimed copies of the step ran before and after the %d timed steps" % args.steps
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # one verified step: counters of this rank's shard (before any further accumulation)
    ctx.reset()
    ctx.query_staged(mode)
    ctx.sync()
    mine = ctx.fetch(mode)
    multi_check = None
    if world > 1:
        # the combined counters on rank 0 must equal the sum of the per-rank results, vector by vector
        vec = np.concatenate([mine["cnt_u"].astype(np.int64), mine["cnt_d"].astype(np.int64),
                              np.array([int(mine["nundet"]), int(mine["nconf"]),
                                        int(mine["rcount_u"].astype(np.uint64).sum()) if mode == cq.MODE_P else 0,
                                        int(mine["rcount_d"].astype(np.uint64).sum()) if mode == cq.MODE_P else 0], dtype=np.int64)])
        tot_t = torch.from_numpy(vec).cuda()
        dist.all_reduce(tot_t)
        combine()
        torch.cuda.synchronize()
        if rank == 0:
            tot = ctx.fetch(mode)
            got = np.concatenate([tot["cnt_u"].astype(np.int64), tot["cnt_d"].astype(np.int64),
                                  np.array([int(tot["nundet"]), int(tot["nconf"]),
                                            int(tot["rcount_u"].astype(np.uint64).sum()) if mode == cq.MODE_P else 0,
                                            int(tot["rcount_d"].astype(np.uint64).sum()) if mode == cq.MODE_P else 0], dtype=np.int64)])
            multi_check = "ok" if np.array_equal(got, tot_t.cpu().numpy()) else "MISMATCH"
        dist.barrier()
    stats = ctx.timing()
    rcount_updates = (int(mine["rcount_u"].astype(np.uint64).sum()) + int(mine["rcount_d"].astype(np.uint64).sum())) if mode == cq.MODE_P else 0
    valid_reads = n - int(mine["n_invalid"])
    hits_per_read = stats["bucket_hits"] / max(valid_reads, 1)
    B = bytes_per_read(rl, h, hits_per_read, rcount_updates / max(n, 1))
    n_steps = max(tm["steps"], 1)
    scan_ms, pack_ms = tm["scan_ms_sum"] / n_steps, tm["pack_ms_sum"] / n_steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n * B / (scan_ms * 1e-3) / 1e9
    prof, prof_warn = profiled_counters(name)
    if prof_warn:
        log("[bench] WARNING: " + prof_warn + " -- roofline.traffic and the ncu-derived fractions are null")
    has_filter = info.filter_bytes > 0
    positions = stats["probes"] / 2
    roof = {"bound": "l2-gather/latency" if has_filter else "hbm-random", "kernel": "scan_reads_kernel",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": prof.get("dram_bytes") if prof else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (streaming copy)" if peaks else "fallback 6650",
            "note": "achieved = ALGORITHMIC bytes (SURVEY 8d: one 32-byte sector per strand and read position) / scan kernel time; "
                    "with the L2-resident filter the probes never reach HBM, so this is not an HBM fraction (it may exceed 1) -- "
                    "the fractions below say what the kernel actually uses",
            "bytes_per_read": B, "scan_ms_per_step": scan_ms, "pack_ms_per_step": pack_ms,
            "probes_per_step": stats["probes"], "candidates_per_read": hits_per_read,
            "chained_loads_per_step": stats["chained_loads"],
            "launch": {"grid": stats["grid_blocks"], "blocks_per_sm": stats["blocks_per_sm"],
                       "dyn_smem": stats["dyn_smem_bytes"], "regs": stats["regs_per_thread"]}}
    if rank == 0:
        fr = {}
        if has_filter:
            # one filter word (L2) per read position: the measured L2 random-gather rate is the roofline
            pct = stats["smem_carveout_pct"]
            smem_block = pct * 228 * 1024 // 100 // max(stats["blocks_per_sm"], 1) - 1024
            g_free = gather_peak(ctx, info.filter_bytes)
            g_kernel = gather_peak(ctx, info.filter_bytes, smem_per_block=smem_block, blocks_per_sm=stats["blocks_per_sm"],
                                   carveout_pct=pct)
            rate = positions / (scan_ms * 1e-3) / 1e9
            fr["l2_gather"] = {"filter_loads_g_per_s": rate, "peak_g_per_s": g_kernel, "frac": rate / g_kernel,
                               "peak_whole_l1_g_per_s": g_free, "smem_carveout_pct": pct,
                               "note": "peak = cq_bench_random_gather over a region of the filter's size, same run, with the scan "
                                       "kernel's shared-memory footprint taken out of each SM's L1 (in-flight gathers live in L1)"}
        else:
            gsec = ctx.bench_random_sectors(1 << 28, iters=2)
            rate = positions / (scan_ms * 1e-3) / 1e9
            fr["hbm_random"] = {"table_sectors_g_per_s": rate, "peak_g_per_s": gsec, "frac": rate / gsec}
        if prof:
            fr["dram"] = {"bytes_per_launch": prof["dram_bytes"], "gb_per_s": prof["dram_bytes"] / (scan_ms * 1e-3) / 1e9,
                          "frac": prof["dram_bytes"] / (scan_ms * 1e-3) / 1e9 / peak}
            fr["issue_slots_busy"] = prof.get("issue_active_pct", 0) / 100.0
            fr["l2_hit_rate"] = prof.get("lts_hit_rate_pct", 0) / 100.0
            fr["ncu_source"] = prof.get("source")
        roof["fractions"] = fr
        gsec = ctx.bench_random_sectors(1 << 28, iters=2)
        roof["random_sector_gather_gsectors_s"] = gsec

    # ---- (2) end to end through the C ABI with HOST buffers ------------------------------------
    pinned_out = ctx.pinned_result_buffers()
    flat = reads.reshape(-1)

    def step_e2e(packed_src=None):
        ctx.reset()
        if world == 1:
            if packed_src is None:
                return ctx.query(mode, flat, None, lengths, stride=rl, buffers=pinned_out)
            return ctx.query_packed(mode, packed_src[0], None, packed_src[1], stride=packed_src[2], buffers=pinned_out)
        # several ranks: the reads go through the same pipeline, the accumulators are reduced on
        # the devices and ONLY the reduced totals cross PCIe, on rank 0
        if packed_src is None:
            ctx.submit(mode, flat, None, lengths, stride=rl)
        else:
            ctx.submit(mode, packed_src[0], None, packed_src[1], stride=packed_src[2], packed=True)
        combine()
        if rank == 0:
            res, keep = ctx._result(mode, 0, False, 0, True, 1 << 16, pinned_out)
            cq.capi._check(cq.lib().cq_fetch(ctx._h, mode, cq.capi.C.byref(res)))
            return ctx._finish(mode, res, keep, 0)
        ctx.sync()
        return None

    def time_e2e(pack_threads, packed_src=None):
        ctx.set_host_packing(pack_threads)
        for _ in range(2):
            step_e2e(packed_src)
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            r = step_e2e(packed_src)
        barrier()
        sec = (time.perf_counter() - t1) / args.steps
        tmq = ctx.timing()
        if world > 1:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, r, tmq

    ncpu = ncpu_unbound
    pack_threads = args.pack_threads if args.pack_threads >= 0 else min(16, ncpu // world)
    if pack_threads < 2 and args.pack_threads < 0:
        pack_threads = 0
    ascii_s, res, tm_ascii = time_e2e(0)
    e2e_paths = {"ascii_over_pcie": {"value": world * n / ascii_s, "ms_per_step": ascii_s * 1e3,
                                     "h2d_bytes_per_step": int(tm_ascii["h2d_bytes"])}}
    e2e_s, h2d, e2e_path = ascii_s, int(tm_ascii["h2d_bytes"]), "ascii_over_pcie"
    if pack_threads > 0:
        pk_s, res, tm_pk = time_e2e(pack_threads)
        e2e_paths["host_packed_2bit"] = {"value": world * n / pk_s, "ms_per_step": pk_s * 1e3,
                                         "h2d_bytes_per_step": int(tm_pk["h2d_bytes"]), "pack_threads": pack_threads,
                                         "host_pack_ms_per_step": tm_pk["host_pack_ms"], "isa": cq.pack_isa()}
        # both are settings of the same C-ABI call (cq_ctx_set_host_packing); the headline is the
        # faster one on this host, the other stays listed under "paths"
        if pk_s < ascii_s:
            e2e_s, h2d, e2e_path = pk_s, int(tm_pk["h2d_bytes"]), "host_packed_2bit"
    # reads packed at the source (what the CLI's FASTQ reader hands over: cq_query_packed): the
    # ASCII never has to be read again, so this path shows what the host-side ceiling costs
    pstride = (rl + 3) // 4
    pk_host = torch.empty((n, pstride), dtype=torch.uint8, pin_memory=True)
    pl_host = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    bad = cq.capi.C.c_uint64()
    cq.capi._check(cq.lib().cq_pack_reads(flat.ctypes.data, None, rl, lengths.ctypes.data, n, max(1, min(16, ncpu // world)),
                                          pk_host.numpy().ctypes.data, pstride, pl_host.numpy().ctypes.data, cq.capi.C.byref(bad)))
    src_s, res_src, tm_src = time_e2e(0, (pk_host.numpy().reshape(-1), pl_host.numpy(), pstride))
    e2e_paths["packed_at_source"] = {"value": world * n / src_s, "ms_per_step": src_s * 1e3,
                                     "h2d_bytes_per_step": int(tm_src["h2d_bytes"]),
                                     "note": "cq_query_packed on reads the caller holds packed (the FASTQ reader's output); "
                                             "reported beside the headline, not as it"}
    e2e_value = world * n / e2e_s
    # the chunked host path must leave exactly the counters of the single resident launch
    if world == 1:
        def same_as_mine(r):
            ok = (int(r["nundet"]), int(r["nconf"])) == (int(mine["nundet"]), int(mine["nconf"]))
            ok = ok and np.array_equal(r["cnt_u"], mine["cnt_u"]) and np.array_equal(r["cnt_d"], mine["cnt_d"])
            if mode == cq.MODE_P:
                ok = ok and np.array_equal(r["rcount_u"], mine["rcount_u"]) and np.array_equal(r["rcount_d"], mine["rcount_d"])
            return ok
        e2e_check = "ok" if same_as_mine(res) and same_as_mine(res_src) else "MISMATCH"
    else:
        e2e_check = None
    d2h = (2 * (G + 1) + 4) * 8 + ((info.n_leaves_u + info.n_leaves_d) * 4 if mode == cq.MODE_P else 0)
    ctx.set_host_packing(-1)

    # ---- (3) CPU baseline beside it (rank 0, N=1 only) + full-vector parity of the sample -----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fq = ensure_sample_fastq(name, w, d)
        threads = host_threads()
        rows = run_reference_harness(d, fq, threads, 2)
        s = w["sample_reads"]
        sample = "first %d of the %d reads, full index, %d OpenMP threads (query64mt_p)" % (s, n, threads)
        if rows:
            r = rows[-1]
            cpu = {"value": s / (r["query_ms"] * 1e-3), "unit": "reads/s", "cores": threads, "kind": "reference",
                   "sample": sample, "index_load_s": r["load_ms"] / 1e3,
                   "note": "std::unordered_map stands in for robin_hood (not vendored by the reference)"}
            ctx.reset()
            g = ctx.query(cq.MODE_P, reads[:s].reshape(-1), None, lengths[:s], stride=rl)
            parity = compare_with_harness(idx, G, g, r, "p")
            ctx.reset()
        else:
            rows = oracle_port_rows(d, w, 1)
            cpu = {"value": s / (rows[-1]["query_ms"] * 1e-3), "unit": "reads/s", "cores": 1, "kind": "port",
                   "sample": "oracle C restatement on %d reads, scaled" % min(s, 50_000)}

    # ---- (4) the other named shapes -------------------------------------------------------------
    secondary = {}
    wanted = [x for x in args.secondary.split(",") if x] if args.secondary else None
    budget_left = lambda: args.time_budget - (time.time() - t_start)  # noqa: E731
    abi_check = None
    if world > 1 and name == "cfg2" and not args.no_secondary:
        # the C++ side's own multi-GPU path (cq_multi_*: sharding + NCCL reduce behind the ABI),
        # exercised on a sample in a process of its own while the other ranks wait on the host
        if rank == 0:
            ns = min(n, 1_000_000)
            single = {"n": ns, "result": {}}
            for m_, tag in ((cq.MODE_P, "p"), (cq.MODE_SC, "sc")):
                ctx.reset()
                r = ctx.query(m_, reads[:ns].reshape(-1), None, lengths[:ns], stride=rl)
                v = [int(r["nundet"]), int(r["nconf"]), [int(x) for x in r["cnt_u"]], [int(x) for x in r["cnt_d"]]]
                if tag == "p":
                    v += [list(rcount_digest(idx, cq.TABLE_U, G, r["rcount_u"])), list(rcount_digest(idx, cq.TABLE_D, G, r["rcount_d"]))]
                else:
                    v += [sorted([a, b, c] for (a, b), c in r["pairs"].items())]
                single["result"][tag] = v
            ctx.reset()
            try:
                abi_check = abi_multi_check(name, w, args.workdir, world, single)
            except Exception as e:
                abi_check = "error: %s" % str(e)[:200]
            log("[bench] abi_multi_gpu_check: %s" % abi_check)
        dist.barrier(group=host_group)

    if not args.no_secondary and name == "cfg2":
        if world == 1 and rank == 0:
            blocks = [("cfg1_refbuilt", lambda: block_cfg1_refbuilt(cq, args.steps, args.warmup)),
                      ("cfg5", lambda: block_synthetic(cq, "cfg5", args.workdir, args.steps, args.warmup, budget_left, modes=("p", "sc"))),
                      ("cfg5_deep", lambda: block_synthetic(cq, "cfg5_deep", args.workdir, args.steps, args.warmup, budget_left)),
                      ("cfg5_dense", lambda: block_synthetic(cq, "cfg5_dense", args.workdir, args.steps, args.warmup, budget_left,
                                                             dedup_ab=True)),
                      ("cfg4", lambda: block_synthetic(cq, "cfg4", args.workdir, max(2, args.steps // 4), 3, budget_left,
                                                       with_harness=False))]
            for bname, fn in blocks:
                if wanted is not None and bname not in wanted:
                    continue
                need = {"cfg4": 200, "cfg1_refbuilt": 20}.get(bname, 60)
                if budget_left() < need:
                    secondary[bname] = {"skipped": "time budget (%.0f s left, block needs about %d s)" % (budget_left(), need)}
                    continue
                tb = time.time()
                try:
                    secondary[bname] = fn()
                except Exception as e:  # a secondary block must not take the headline down with it
                    secondary[bname] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
                secondary[bname]["block_s"] = time.time() - tb
                log("[bench] secondary %s: %s" % (bname, json.dumps(secondary[bname])[:600]))
        elif world > 1 and (wanted is None or "cfg3_sc_strong" in wanted):
            # configs[2] as specified: query64_sc path, 150-bp reads, 50M reads sharded over the ranks
            w3 = dict(WORKLOADS["cfg3"])
            total = STRONG_CFG3_READS if not args.reads else args.reads * world
            lo, hi = multigpu.shard_range(total, rank, world)
            n3, rl3 = hi - lo, w3["read_len"]
            del host, reads
            r3 = make_reads(w3, lo, n3)
            l3 = np.full(n3, rl3, dtype=np.uint8)
            ctx.stage(r3.reshape(-1), None, l3, stride=rl3)
            old_mode, mode = mode, cq.MODE_SC
            for _ in range(3):
                step_resident()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k3 = max(3, args.steps // 2)
            a0.record()
            for _ in range(k3):
                step_resident()
            join_reduces()
            a1.record()
            barrier()
            t = torch.tensor([a0.elapsed_time(a1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms3 = float(t.item()) / k3
            # parity: the reduced totals == the sum of the ranks' own results; rank 0's sample == the reference
            ctx.reset()
            ctx.query_staged(cq.MODE_SC)
            ctx.sync()
            own = ctx.fetch(cq.MODE_SC)
            pm = multigpu.gather_pair_maps(own["pairs"])
            vec = torch.from_numpy(np.concatenate([own["cnt_u"].astype(np.int64), own["cnt_d"].astype(np.int64),
                                                   np.array([int(own["nundet"]), int(own["nconf"])], dtype=np.int64)])).cuda()
            dist.all_reduce(vec)
            combine()
            torch.cuda.synchronize()
            blk = None
            if rank == 0:
                tot = ctx.fetch(cq.MODE_SC)
                got = np.concatenate([tot["cnt_u"].astype(np.int64), tot["cnt_d"].astype(np.int64),
                                      np.array([int(tot["nundet"]), int(tot["nconf"])], dtype=np.int64)])
                ok = np.array_equal(got, vec.cpu().numpy())
                fq = ensure_sample_fastq("cfg3", w3, d)
                rows = run_reference_harness(d, fq, 1, 1, mode="sc")
                ref_v = None
                if rows:
                    s3 = w3["sample_reads"]
                    ctx.reset()
                    g = ctx.query(cq.MODE_SC, r3[:s3].reshape(-1), None, l3[:s3], stride=rl3)
                    ref_v = compare_with_harness(idx, G, g, rows[-1], "sc")
                blk = {"workload": "BASELINE configs[2]: cfg2 index, query64_sc path, %d x %dbp reads sharded over %d GPUs (strong scaling), "
                                   "one NCCL reduce of the counter block per step, pair maps merged on rank 0" % (total, rl3, world),
                       "reads_per_s": total / (ms3 * 1e-3), "ms_per_step": ms3, "scaling": "strong", "reads_per_gpu": n3,
                       "reduced_totals_equal_sum_of_ranks": bool(ok), "distinct_pairs": len(pm),
                       "reference_sample_query64_sc": ref_v,
                       "cpu_reference_reads_per_s_query64_sc_1_thread": (w3["sample_reads"] / (rows[-1]["query_ms"] * 1e-3)) if rows else None,
                       "parity": "ok" if ok and ref_v in (None, "ok") else "MISMATCH"}
            mode = old_mode
            dist.barrier()
            if rank == 0:
                secondary["cfg3_sc_strong"] = blk

    if rank == 0:
        line = {
            "metric": "reads/sec classified", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": name + ": " + w["desc"], "reads_per_gpu_per_step": n, "read_len": rl,
                       "hash_len": h, "mode": "query64_p" if mode == cq.MODE_P else "query64_sc",
                       "leaves_u": info.n_leaves_u, "leaves_d": info.n_leaves_d,
                       "table_gb": info.n_table_buckets * 32 / 1e9, "index_device_gb": info.device_bytes / 1e9,
                       "filter_mb": info.filter_bytes / (1 << 20),
                       "l2_policy": "inputs larger than L2: %.2f GB prefix table + %.2f GB reads per step" % (
                           info.n_table_buckets * 32 / 1e9, n * rl / 1e9),
                       "parallelism": ("index replicated, reads sharded, counters combined over NCCL once per step (counter block + one grouped launch for the two rcount arrays)"
                                       + ("; the reduce of step k runs on a side stream while step k+1 scans into the context's other accumulator set (cq_swap_accumulators)" if overlap else "")) if world > 1 else "1 GPU",
                       "index_prepare_s": t_index, "numa": numa},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "d2h_note": "reduced totals, rank 0 only" if world > 1 else "totals",
                    "ms_per_step": e2e_s * 1e3, "path": e2e_path, "paths": e2e_paths},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity_vs_reference_sample": parity,
            "parity_compared": "nundet, nconf, cnt_u[1..G], cnt_d[1..G], per-leaf rcount sums and position-weighted digests over "
                               "Hash::map_sp (both tables) -- ref_harness query64mt_p on the sample",
            "e2e_equals_resident_launch": e2e_check,
            "multi_gpu_reduce_check": multi_check,
            "abi_multi_gpu_check": abi_check,
            "secondary": secondary,
            "result": {"nundet": int(mine["nundet"]), "nconf": int(mine["nconf"]),
                       "sum_u": int(mine["cnt_u"].sum()), "sum_d": int(mine["cnt_d"].sum())},
            "bench_wall_s": time.time() - t_start,
        }
        emit(json_fd, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
