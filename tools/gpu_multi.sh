#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-GPU tests, then both arms of bench.py under torchrun.
n=${1:-2}
tag=${2:-multi$n}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q > $out/gputests.log 2>&1
echo "tests rc=$?"; tail -3 $out/gputests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; tail -c 1500 $out/bench.err; cat $out/bench.json | cut -c1-3000
