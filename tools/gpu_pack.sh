#!/bin/bash
# pack kernel A/B (current build against the pv1 variant), its tests, one ncu capture, host packer thread counts
tag=${1:-pack}
out=gpurun_out/$tag
mkdir -p $out
nproc > $out/nproc.txt; lscpu | head -25 >> $out/nproc.txt
for rep in 1 2; do
  python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
  python tools/kernel_ab.py --packed >> $out/ab.jsonl 2>> $out/ab.err
  for v in cammiq_b200/variants/*.so; do
    CAMMIQ_LIB=$PWD/$v python tools/kernel_ab.py >> $out/ab.jsonl 2>> $out/ab.err
    CAMMIQ_LIB=$PWD/$v python tools/kernel_ab.py --packed >> $out/ab.jsonl 2>> $out/ab.err
  done
done
python - <<P
import json
for l in open("$out/ab.jsonl"):
    d=json.loads(l); print(d['lib'], 'packed' if d['packed'] else 'ascii', 'scan %.3f pack %.3f'%(d['scan_ms_mean'],d['pack_ms_mean']), d['checksum'][:4])
P
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_packed.py -m gpu -x -q > $out/tests.log 2>&1; tail -3 $out/tests.log
ncu --set full --import-source on --clock-control none -k regex:pack_tiles -c 1 -o $out/pack_ascii python tools/kernel_ab.py --iters 1 > $out/ncu1.log 2>&1
ncu -i $out/pack_ascii.ncu-rep --page details > $out/pack_ascii_details.txt 2>&1
for t in 16 24 30; do
  python bench.py --no-secondary --pack-threads $t > $out/bench_t$t.json 2> $out/bench_t$t.err
  python - <<P
import json
d=json.loads(open("$out/bench_t$t.json").read().strip().splitlines()[-1])
print($t, d['value'], d['ms_per_step'], d['e2e']['paths'].get('host_packed_2bit'))
P
done
