#!/bin/bash
# One GPU-box session: A/B of the scan-kernel builds, the GPU tests, one ncu capture.
# Everything lands in gpurun_out/<tag>/.
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
python tools/kernel_ab.py > $out/ab_main.json 2> $out/ab_main.err
echo "main rc=$?"; cat $out/ab_main.json
timeout 1200 python -m pytest tests -m gpu -x -q > $out/gputests.log 2>&1
echo "tests rc=$?"; tail -5 $out/gputests.log
for v in none; do
  CAMMIQ_LIB=$PWD/cammiq_b200/variants/libcammiq_gpu_$v.so python tools/kernel_ab.py >> $out/ab_variants.jsonl 2>> $out/ab_variants.err
done
cat $out/ab_variants.jsonl
python tools/kernel_ab.py --packed >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --workload cfg3 >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --filter-mb 0 >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --filter-mb 8 >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --filter-mb 16 >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --filter-mb 48 >> $out/ab_more.jsonl 2>> $out/ab_more.err
python tools/kernel_ab.py --mode sc >> $out/ab_more.jsonl 2>> $out/ab_more.err
cat $out/ab_more.jsonl
ncu --set full --import-source on --clock-control none -k regex:scan_reads -c 1 -o $out/scan_full python tools/kernel_ab.py --iters 1 > $out/ncu.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:scan_reads -c 1 --csv --log-file $out/ncu_metrics.csv python tools/kernel_ab.py --iters 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $out/ncu_bench.log 2>&1
ls -la $out
