#!/bin/bash
tag=${1:-exp}
out=gpurun_out/$tag
mkdir -p $out
python tools/microbench_l1.py > $out/l1.log 2>&1; cp gpurun_out/microbench_l1.json $out/ 2>/dev/null
cat $out/l1.log
python tools/kernel_ab.py > $out/ab.jsonl 2>$out/ab.err
python tools/kernel_ab.py --packed >> $out/ab.jsonl 2>>$out/ab.err
cat $out/ab.jsonl
ncu --set full --import-source on --clock-control none -k regex:scan_reads -c 1 -o $out/scan_ascii python tools/kernel_ab.py --iters 1 > $out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_reads -c 1 -o $out/scan_packed python tools/kernel_ab.py --iters 1 --packed > $out/ncu2.log 2>&1
ls -la $out
