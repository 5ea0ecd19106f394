#!/bin/bash
# Lean end-of-round session: A/B of the tile hand-out, GPU tests, ncu figures stamped before the bench reads them,
# the bench line, smoke.  Everything lands in gpurun_out/<tag>/.
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
for v in "" cammiq_b200/variants/*.so; do
  CAMMIQ_LIB=${v:+$PWD/$v} python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
  CAMMIQ_LIB=${v:+$PWD/$v} python tools/kernel_ab.py --workload cfg5 >> $out/ab.jsonl 2>> $out/ab.err
done
python tools/kernel_ab.py --workload cfg3 >> $out/ab.jsonl 2>> $out/ab.err
python tools/kernel_ab.py --random-reads >> $out/ab.jsonl 2>> $out/ab.err
python tools/kernel_ab.py --mode sc >> $out/ab.jsonl 2>> $out/ab.err
python - <<P
import json
for l in open("$out/ab.jsonl"):
    d=json.loads(l); print(d['lib'], d['workload'], 'rand' if d['random_reads'] else '', 'scan %.3f pack %.3f'%(d['scan_ms_mean'],d['pack_ms_mean']), 'regs',d['regs'], d['checksum'][:4])
P
timeout 1500 python -m pytest tests -m gpu -x -q > $out/gputests.log 2>&1
echo "tests rc=$?"; tail -3 $out/gputests.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
ncu --metrics $M --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics_cfg2.csv python tools/kernel_ab.py --iters 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics_cfg3.csv python tools/kernel_ab.py --iters 1 --workload cfg3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics_cfg2_random_reads.csv python tools/kernel_ab.py --iters 1 --random-reads > /dev/null 2>&1
python tools/ncu_to_traffic.py $out/ncu_metrics_cfg2.csv cfg2 profiles/r02_scan_ncu_metrics_cfg2.csv > $out/stamp.log 2>&1
python tools/ncu_to_traffic.py $out/ncu_metrics_cfg3.csv cfg3 profiles/r02_scan_ncu_metrics_cfg3.csv >> $out/stamp.log 2>&1
python tools/ncu_to_traffic.py $out/ncu_metrics_cfg2_random_reads.csv cfg2_random_reads profiles/r02_scan_ncu_metrics_cfg2_random_reads.csv >> $out/stamp.log 2>&1
cp profiles/scan_traffic.json $out/scan_traffic.json
python bench.py > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; tail -c 300 $out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $out/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
