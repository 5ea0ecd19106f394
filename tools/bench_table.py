#!/usr/bin/env python
"""Markdown summary of a bench.py JSON line (and optionally the reference arm's):
    python tools/bench_table.py gpurun_out/round1/bench.json [gpurun_out/round1/bench_reference.json]"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ref = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]) if len(sys.argv) > 2 else None
r = d["roofline"]
print("| headline (%s) | value |" % d["config"]["workload"].split(":")[0])
print("|---|---|")
print("| `value`: reads resident in HBM, step = reset + pack + scan + count reduction | **%.3g reads/s** (%.2f ms per %d reads: pack %.2f + scan %.2f) |" % (
    d["value"], d["ms_per_step"], d["config"]["reads_per_gpu_per_step"], r["pack_ms_per_step"], r["scan_ms_per_step"]))
print("| `e2e`: `cq_query`, host buffers in, counters + rcount back (%s) | **%.3g reads/s** (%.2f ms) |" % (d["e2e"]["path"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
for k, v in d["e2e"]["paths"].items():
    print("| e2e path `%s` | %.3g reads/s (%.2f ms, %.0f MB H2D) |" % (k, v["value"], v["ms_per_step"], v["h2d_bytes_per_step"] / 1e6))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    print("| reference `query64mt_p`, %d host threads, same run | %.3g reads/s |" % (c["cores"], c["value"]))
if ref:
    print("| reference arm (`--impl reference`) | %.3g reads/s |" % ref["value"])
f = r.get("fractions", {})
if "l2_gather" in f:
    g = f["l2_gather"]
    print("| roofline: L2 gather | %.0f of %.0f G filter loads/s = **%.2f** (carve-out %s %%) |" % (g["filter_loads_g_per_s"], g["peak_g_per_s"], g["frac"], g.get("smem_carveout_pct")))
if "dram" in f:
    print("| roofline: DRAM | %.1f GB per launch = %.0f GB/s = **%.2f** of %.0f |" % (f["dram"]["bytes_per_launch"] / 1e9, f["dram"]["gb_per_s"], f["dram"]["frac"], r["peak"]))
    print("| issue slots busy / L2 hit rate (ncu) | %.2f / %.2f |" % (f["issue_slots_busy"], f["l2_hit_rate"]))
print("| algorithmic bytes (SURVEY 8d) / HBM peak | %.0f GB/s / %.0f = %.2f (probes answered by L2: not an HBM fraction) |" % (r["achieved"], r["peak"], r["frac"]))
print("| parity (vectors + rcount digests vs reference, 200k reads) | %s; chunked host path == resident launch: %s |" % (d["parity_vs_reference_sample"], d["e2e_equals_resident_launch"]))
print()
print("| secondary block | reads/s | ms/step | parity | notes |")
print("|---|---|---|---|---|")
for k, v in d.get("secondary", {}).items():
    if "reads_per_s" not in v:
        print("| %s | — | — | — | %s |" % (k, v.get("skipped") or v.get("error")))
        continue
    notes = []
    if "roofline" in v:
        notes.append("%s: %.1f of %.1f %s = %.2f" % (v["roofline"]["bound"], v["roofline"]["achieved"], v["roofline"]["peak"], v["roofline"]["unit"], v["roofline"]["frac"]))
    if "reads_per_s_sc" in v:
        notes.append("query64_sc %.3g reads/s" % v["reads_per_s_sc"])
    if "cpu_reference_reads_per_s" in v:
        notes.append("reference CPU " + ", ".join("%s %.3g" % kv for kv in v["cpu_reference_reads_per_s"].items()))
    if "candidates_per_read" in v:
        notes.append("%.1f candidates, %.1f leaf hits per read" % (v["candidates_per_read"], v["leaf_hits_per_read"]))
    print("| %s | %.3g | %.2f | %s | %s |" % (k, v["reads_per_s"], v["ms_per_step"], v.get("parity"), "; ".join(notes)))
