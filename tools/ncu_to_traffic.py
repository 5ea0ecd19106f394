#!/usr/bin/env python
"""ncu metrics CSV of the scan kernel -> profiles/scan_traffic.json.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,\
lts__t_sectors.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum \
        --clock-control none -k regex:scan_reads -c 1 --csv --log-file gpurun_out/x/ncu_metrics.csv \
        python tools/kernel_ab.py --iters 1
    python tools/ncu_to_traffic.py gpurun_out/x/ncu_metrics.csv cfg2 profiles/r02_scan_ncu_metrics.csv

The JSON is stamped with the git blob hash of cammiq_b200/csrc/scan_kernels.cuh as it is NOW (run
this right after the capture, before editing the kernel): bench.py reports the figures only while
the kernel source still matches, and `traffic: null` with a warning otherwise."""
import csv
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def blob_hash(path):
    data = open(path, "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def main():
    src, workload, keep = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    name, val, kern = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    m = {}
    for r in rows[1:]:
        m[r[name]] = float(r[val].replace(",", ""))
        kernel = r[kern]
    out_path = os.path.join(REPO, "profiles", "scan_traffic.json")
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    blob = blob_hash(os.path.join(REPO, "cammiq_b200", "csrc", "scan_kernels.cuh"))
    if out.get("scan_kernels_cuh_blob") != blob:
        out = {}   # figures of another kernel source are void
    shutil.copy(src, os.path.join(REPO, keep))
    out["scan_kernels_cuh_blob"] = blob
    out["note"] = ("ncu figures of scan_reads_kernel per launch, valid for the scan_kernels.cuh whose git blob hash is "
                   "scan_kernels_cuh_blob; written by tools/ncu_to_traffic.py")
    out[workload] = {
        "kernel": kernel,
        "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
        "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
        "gpu_time_ms": m["gpu__time_duration.sum"] / 1e6,
        "lts_hit_rate_pct": m.get("lts__t_sector_hit_rate.pct"),
        "lts_sectors": m.get("lts__t_sectors.sum"),
        "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "executed_warp_instructions": m.get("smsp__inst_executed.sum"),
        "global_load_requests": m.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
        "source": keep,
    }
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out[workload], indent=1))


if __name__ == "__main__":
    main()
