#!/bin/bash
# host packer A/B on one box: the e2e paths of bench.py with and without the streaming conversion
out=gpurun_out/${1:-e2e}
mkdir -p $out
for rep in 1 2; do
for s in 1 0; do
CAMMIQ_PACK_STREAM=$s python bench.py --no-secondary --no-cpu-baseline > $out/bench_s${s}_$rep.json 2> $out/bench_s${s}_$rep.err
python - <<P
import json
d=json.loads(open("$out/bench_s${s}_$rep.json").read().strip().splitlines()[-1])
p=d["e2e"]["paths"]["host_packed_2bit"]
print("stream=$s rep $rep: e2e %.2f ms, host pack %.2f ms, equal=%s, value %.3g"%(p["ms_per_step"], p["host_pack_ms_per_step"], d["e2e_equals_resident_launch"], d["value"]))
P
done
done
