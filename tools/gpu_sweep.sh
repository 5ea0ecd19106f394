#!/bin/bash
# filter-size sweep of the scan kernel (cfg2, resident reads) + ncu figures of the default build
out=gpurun_out/${1:-sweep}
mkdir -p $out
for mb in 24 32 40 48 56 64; do
  python tools/kernel_ab.py --filter-mb $mb >> $out/sweep.jsonl 2>> $out/sweep.err
done
python tools/kernel_ab.py --workload cfg3 --filter-mb 48 >> $out/sweep.jsonl 2>> $out/sweep.err
cat $out/sweep.jsonl
ncu --set full --import-source on --clock-control none -k regex:scan_reads -c 1 -o $out/scan_full python tools/kernel_ab.py --iters 1 > $out/ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics.csv python tools/kernel_ab.py --iters 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics_f40.csv python tools/kernel_ab.py --iters 1 --filter-mb 40 > /dev/null 2>&1
