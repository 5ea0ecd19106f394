#!/bin/bash
# A/B of the experiment builds in cammiq_b200/variants (cfg2, resident reads, mode p unless said otherwise)
out=gpurun_out/${1:-variants}
mkdir -p $out
python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
for v in $(ls cammiq_b200/variants/*.so); do
  CAMMIQ_LIB=$PWD/$v python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
done
python tools/kernel_ab.py --mode sc >> $out/ab.jsonl 2>> $out/ab.err
cat $out/ab.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['lib'], 'scan %.3f pack %.3f'%(d['scan_ms_mean'],d['pack_ms_mean']), 'regs',d['regs'],'bps',d['blocks_per_sm'], d['checksum'][:4])
"
