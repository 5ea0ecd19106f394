#!/bin/bash
# e2e paths of bench.py at several pipeline chunk sizes, same box
out=gpurun_out/${1:-e2e}
mkdir -p $out
for c in 1048576 524288 262144 1048576 524288; do
CAMMIQ_CHUNK_READS=$c python bench.py --no-secondary --no-cpu-baseline > $out/bench_c$c.json 2> $out/bench_c$c.err
python - <<P
import json
d=json.loads(open("$out/bench_c$c.json").read().strip().splitlines()[-1])
p=d["e2e"]["paths"]
print("chunk $c: host-packed %.2f ms (pack %.2f), ascii %.2f ms, packed-at-source %.2f ms, equal=%s"%(p["host_packed_2bit"]["ms_per_step"], p["host_packed_2bit"]["host_pack_ms_per_step"], p["ascii_over_pcie"]["ms_per_step"], p["packed_at_source"]["ms_per_step"], d["e2e_equals_resident_launch"]))
P
done
