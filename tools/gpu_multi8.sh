#!/bin/bash
# bench.py under torchrun on 8 GPUs; extra arguments go to bench.py (e.g. --no-secondary)
out=gpurun_out/${1:-multi8}
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
n=8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 "${@:2}" > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "bench n=$n rc=$?"; tail -c 600 $out/bench_n$n.err; cut -c1-1500 $out/bench_n$n.json
