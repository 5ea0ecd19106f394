#!/usr/bin/env python
"""bench.py -- reads/s of the CAMMiQ read-matching hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU scan

A step = one pass of the hot path (pack -> scan -> count reduction [-> NCCL reduce]) over one
batch of synthetic reads.  Workload at N=1 = BASELINE.json configs[1]: `--both` index over
500 synthetic strain genomes (~1.5 Gbp), 10M simulated 100-bp reads with 1% substitutions.
N>1: index replicated per GPU, every rank scans its own 10M reads (weak scaling), counters
combined with one NCCL reduce per step.  One JSON line on stdout (rank 0).
"""
import argparse
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
    # BASELINE.json configs[0] (the reference's own CPU-runnable case); parity/smoke size
    "cfg1": dict(n_genomes=10, genome_len=1_000_000, cluster_size=3, reads=100_000, read_len=100,
                 erate=0.0, k=26, lmax=50, seed=1, sample_reads=100_000,
                 desc="10 synthetic 1 Mbp genomes, 100k error-free 100bp reads"),
    # configs[2]-like shape: 150 bp reads on the cfg2 index
    "cfg3": dict(n_genomes=500, genome_len=3_000_000, cluster_size=4, reads=10_000_000, read_len=150,
                 erate=0.01, k=26, lmax=50, seed=2, sample_reads=200_000,
                 desc="cfg2 index, 150bp reads"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_params(w, threads=0):
    from cammiq_b200 import synthlib as sl
    return sl.params(seed=w["seed"], n_genomes=w["n_genomes"], genome_len=w["genome_len"],
                     cluster_size=w["cluster_size"], k=w["k"], lmax=w["lmax"], threads=threads,
                     permille_deep=w.get("permille_deep", 50))


def workdir_for(name, w, base):
    return os.path.join(base, "cammiq_bench_%s_g%d_l%d_s%d_d%d" % (name, w["n_genomes"], w["genome_len"], w["seed"],
                                                                    w.get("permille_deep", 50)))


def ensure_index(name, w, base):
    """Synthetic index files (format-exact .bin1/.bin2 + map + meta), generated once."""
    from cammiq_b200 import synthlib as sl
    d = workdir_for(name, w, base)
    done = os.path.join(d, ".done")
    if not os.path.exists(done):
        t = time.time()
        st = sl.write_index(synth_params(w), d)
        open(done, "w").write(json.dumps(st))
        log("[bench] synthetic index written to %s in %.1f s: %s" % (d, time.time() - t, st))
    return d


def ensure_sample_fastq(name, w, d):
    from cammiq_b200 import synthlib as sl
    fq = os.path.join(d, "sample_%d_%d.fq" % (w["sample_reads"], w["read_len"]))
    if not os.path.exists(fq):
        sl.write_fastq(synth_params(w), 0, w["sample_reads"], w["read_len"], w["erate"], fq + ".tmp")
        os.rename(fq + ".tmp", fq)
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


def run_reference_harness(d, fq, threads, reps, mode="mt"):
    """The UNMODIFIED reference scan (oracle/_ref/ref_harness) on the host cores."""
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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
    from cammiq_b200 import synthlib as sl
    n = min(w["sample_reads"], 50_000)
    reads = sl.make_reads(synth_params(w), 0, n, w["read_len"], w["erate"])
    ou, od = ol.OracleIndex(os.path.join(d, "index_u.bin1")), ol.OracleIndex(os.path.join(d, "index_d.bin2"))
    lens = np.full(n, w["read_len"], np.uint8)
    offs = np.arange(n, dtype=np.uint64) * w["read_len"]
    rows = []
    for _ in range(reps):
        t = time.time()
        ol.oracle_query(ou, od, ol.MODE_P, w["n_genomes"], reads.reshape(-1), offs, lens)
        rows.append({"query_ms": (time.time() - t) * 1e3 * (w["sample_reads"] / n), "load_ms": 0})
    return rows


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
    ap.add_argument("--mode", default="p", choices=["p", "sc"])
    ap.add_argument("--deep-permille", type=int, default=-1, help="override the share of keys longer than h (synthetic index)")
    ap.add_argument("--pack-threads", type=int, default=-1,
                    help="host threads packing reads to 2 bits before the copy in the e2e leg (0 = ASCII over PCIe, "
                         "-1 = min(16, host cpus / ranks))")
    ap.add_argument("--filter-mb", type=float, default=-1, help="override the membership-filter budget (MB, 0 = none)")
    args = ap.parse_args()
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
    from cammiq_b200 import synthlib as sl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback.")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = cq.MODE_P if args.mode == "p" else cq.MODE_SC

    # ---- workload: index (rank 0 writes, everyone loads), reads (each rank its own shard) ----
    t0 = time.time()
    if rank == 0:
        d = ensure_index(name, w, args.workdir)
    if world > 1:
        dist.barrier()
    d = workdir_for(name, w, args.workdir)
    idx = cq.Index(os.path.join(d, "index_u.bin1"), os.path.join(d, "index_d.bin2"))
    if args.filter_mb >= 0:
        idx.set_filter_budget(int(args.filter_mb * (1 << 20)))
    info = idx.info
    # everything (kernels, memsets, NCCL hand-off, the timing events) runs on one torch stream
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream().cuda_stream
    ctx = cq.Context(local, stream=stream).upload(idx, w["n_genomes"])
    t_index = time.time() - t0
    n, rl, h = w["reads"], w["read_len"], info.hash_len
    host = torch.empty((n, rl), dtype=torch.uint8, pin_memory=True)
    reads = host.numpy()
    sl.make_reads(synth_params(w), rank * n, n, rl, w["erate"], out=reads)
    lengths_t = torch.full((n,), rl, dtype=torch.uint8).pin_memory()
    lengths = lengths_t.numpy()
    log("[bench] rank %d: index %d U + %d D leaves, table %.2f GB, %d reads ready (%.1f s)" % (
        rank, info.n_leaves_u, info.n_leaves_d, info.n_table_buckets * 32 / 1e9, n, time.time() - t0))

    counts, rc_u, rc_d = (torch.as_tensor(a, device="cuda") for a in ctx.device_counter_arrays())

    def combine():
        # the one collective of the path: sum the counter block (and the per-leaf rcount in
        # mode P) into rank 0 over NCCL
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
        ctx.reset()
        ctx.query_staged(mode)
        combine()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    ctx.timing_reset()
    launches0 = ctx.timing()["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    tm = ctx.timing()
    clocks = sampler.stop()
    launches = tm["kernel_launches"] - launches0
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
        # the combined counters on rank 0 must equal the sum of the per-rank results
        local = torch.tensor([int(mine["nundet"]), int(mine["nconf"]), int(mine["cnt_u"].sum()),
                              int(mine["cnt_d"].sum()), int(mine["rcount_u"].sum()) if mode == cq.MODE_P else 0],
                             device="cuda", dtype=torch.int64)
        dist.all_reduce(local)
        combine()
        torch.cuda.synchronize()
        if rank == 0:
            tot = ctx.fetch(mode)
            got = [int(tot["nundet"]), int(tot["nconf"]), int(tot["cnt_u"].sum()), int(tot["cnt_d"].sum()),
                   int(tot["rcount_u"].sum()) if mode == cq.MODE_P else 0]
            multi_check = "ok" if got == local.tolist() else "MISMATCH %s vs %s" % (got, local.tolist())
        dist.barrier()
    stats = ctx.timing()
    rcount_updates = (int(mine["rcount_u"].sum()) + int(mine["rcount_d"].sum())) if mode == cq.MODE_P else 0
    valid_reads = n - int(mine["n_invalid"])
    hits_per_read = stats["bucket_hits"] / max(valid_reads, 1)
    B = bytes_per_read(rl, h, hits_per_read, rcount_updates / max(n, 1))
    scan_ms = tm["scan_ms_sum"] / max(tm["steps"], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n * B / (scan_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "scan_traffic.json"))).get(name)
    except Exception:
        pass
    gsec = ctx.bench_random_sectors(1 << 28, iters=2) if rank == 0 else 0.0

    # ---- (2) end to end through the C ABI with HOST buffers ------------------------------------
    pinned_out = ctx.pinned_result_buffers()

    def step_e2e():
        ctx.reset()
        r = ctx.query(mode, reads.reshape(-1), None, lengths, stride=rl, buffers=pinned_out)
        combine()
        return r

    def time_e2e(pack_threads):
        ctx.set_host_packing(pack_threads)
        for _ in range(2):
            step_e2e()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            r = step_e2e()
        barrier()
        sec = (time.perf_counter() - t1) / args.steps
        tmq = ctx.timing()
        if world > 1:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, r, tmq

    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
    e2e_value = world * n / e2e_s
    # the chunked host path must leave exactly the counters of the single resident launch
    if world == 1:
        same = (int(res["nundet"]), int(res["nconf"]), int(res["cnt_u"].sum()), int(res["cnt_d"].sum())) == (
            int(mine["nundet"]), int(mine["nconf"]), int(mine["cnt_u"].sum()), int(mine["cnt_d"].sum()))
        if mode == cq.MODE_P:
            same = same and np.array_equal(res["rcount_u"], mine["rcount_u"]) and np.array_equal(res["rcount_d"], mine["rcount_d"])
        e2e_check = "ok" if same else "MISMATCH"
    else:
        e2e_check = None
    d2h = (2 * (w["n_genomes"] + 1) + 4) * 8 + ((info.n_leaves_u + info.n_leaves_d) * 4 if mode == cq.MODE_P else 0)
    ctx.set_host_packing(-1)

    # ---- (3) CPU baseline beside it (rank 0, N=1 only) + parity of the sample -------------------
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
            mine_t = (int(g["nundet"]), int(g["nconf"]), int(g["cnt_u"].sum()), int(g["cnt_d"].sum()))
            ref_t = (r["nundet"], r["nconf"], r["sum_u"], r["sum_d"])
            parity = "ok" if mine_t == ref_t else "MISMATCH gpu=%s ref=%s" % (mine_t, ref_t)
        else:
            rows = oracle_port_rows(d, w, 1)
            cpu = {"value": s / (rows[-1]["query_ms"] * 1e-3), "unit": "reads/s", "cores": 1, "kind": "port",
                   "sample": "oracle C restatement on %d reads, scaled" % min(s, 50_000)}

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
                       "parallelism": "index replicated, reads sharded, 1 NCCL reduce/step" if world > 1 else "1 GPU",
                       "index_prepare_s": t_index},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "path": e2e_path, "paths": e2e_paths},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "scan_reads_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (streaming copy)" if peaks else "fallback 6650",
                         "bytes_per_read": B, "scan_ms_per_step": scan_ms,
                         "pack_ms_per_step": tm["pack_ms_sum"] / max(tm["steps"], 1),
                         "probes_per_step": stats["probes"], "bucket_hits_per_read": hits_per_read,
                         "chained_loads_per_step": stats["chained_loads"],
                         "launch": {"grid": stats["grid_blocks"], "blocks_per_sm": stats["blocks_per_sm"],
                                    "dyn_smem": stats["dyn_smem_bytes"], "regs": stats["regs_per_thread"]},
                         "probe_rate_gprobes_s": stats["probes"] / (scan_ms * 1e-3) / 1e9,
                         "random_sector_gather_gsectors_s": gsec,
                         "frac_of_random_gather": (stats["probes"] / (scan_ms * 1e-3) / 1e9) / gsec if gsec else None},
            "cpu_baseline": cpu,
            "parity_vs_reference_sample": parity,
            "e2e_equals_resident_launch": e2e_check,
            "multi_gpu_reduce_check": multi_check,
            "result": {"nundet": int(mine["nundet"]), "nconf": int(mine["nconf"]),
                       "sum_u": int(mine["cnt_u"].sum()), "sum_d": int(mine["cnt_d"].sum())},
        }
        emit(json_fd, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
