#!/bin/bash
# 8-GPU bench: dynamic hand-out across 8 ranks + one NCCL all-reduce
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpus8.txt
echo "=== bench 8 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8.log 2>&1; grep '^{' gpurun_out/bench_n8.log | cut -c1-1500; tail -3 gpurun_out/bench_n8.log | cut -c1-300
