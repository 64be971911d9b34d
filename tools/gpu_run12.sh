#!/bin/bash
# 2-GPU bench (dynamic hand-out + NCCL all-reduce) and its reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpus2.txt
echo "=== bench 2 gpus dynamic"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-1200
echo "=== bench 2 gpus static"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --static > gpurun_out/bench_n2_static.log 2>&1; tail -1 gpurun_out/bench_n2_static.log | cut -c1-600
echo "=== bench ref 2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2>&1; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
