#!/bin/bash
out=gpurun_out/${1:-bench1}
mkdir -p $out
python bench.py > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; tail -c 300 $out/bench.err
python - <<P
import json
d=json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["clocks"], d["e2e"]["value"], d["roofline"]["traffic"], d["bench_wall_s"])
P
