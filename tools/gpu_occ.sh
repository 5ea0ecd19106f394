#!/bin/bash
# occupancy variants of the scan kernel (warps per CTA x CTAs per SM) against the current build
tag=${1:-occ}
out=gpurun_out/$tag
mkdir -p $out
for rep in 1 2; do
  python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
  for v in cammiq_b200/variants/*.so; do
    CAMMIQ_LIB=$PWD/$v python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
  done
done
for v in "" cammiq_b200/variants/*.so; do
  CAMMIQ_LIB=${v:+$PWD/$v} python tools/kernel_ab.py --workload cfg5 >> $out/ab.jsonl 2>> $out/ab.err
  CAMMIQ_LIB=${v:+$PWD/$v} python tools/kernel_ab.py --random-reads >> $out/ab.jsonl 2>> $out/ab.err
done
python - <<P
import json
for l in open("$out/ab.jsonl"):
    d=json.loads(l); print(d['lib'], d['workload'], 'rand' if d['random_reads'] else '', 'scan %.3f pack %.3f'%(d['scan_ms_mean'],d['pack_ms_mean']), 'regs',d['regs'],'bps',d['blocks_per_sm'],'smem',d['dyn_smem'], d['checksum'][:4])
P
