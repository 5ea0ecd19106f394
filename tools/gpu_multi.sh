#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the multi-GPU tests, then bench.py under torchrun with and without the
# overlapped reduce (headline only), then the full bench line.
n=${1:-2}
tag=${2:-multi$n}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py tests/test_gpu_edges.py -m gpu -q > $out/gputests.log 2>&1
echo "tests rc=$?"; tail -3 $out/gputests.log
run() { timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 10 --warmup 3 "${@:3}" > $out/$2.json 2> $out/$2.err; echo "$2 rc=$?"; tail -c 400 $out/$2.err; python - <<P
import json
try:
    d=json.loads(open("$out/$2.json").read().strip().splitlines()[-1])
    print("$2", d["value"], d["ms_per_step"], d.get("multi_gpu_check"), {k: round(v["value"]/1e9,3) for k,v in d["e2e"]["paths"].items()})
except Exception as e:
    print("$2: no line", e)
P
}
CAMMIQ_NO_REDUCE_OVERLAP=1 run 29511 bench_serial --no-secondary
run 29512 bench_overlap --no-secondary
run 29513 bench
cut -c1-2500 $out/bench.json
