#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
export SWEEP_CONFIGS="1,1,1"
echo "=== sweep benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 14 27 > gpurun_out/sweep_benzene.log 2>&1; cat gpurun_out/sweep_benzene.log
echo "=== sweep 60/500 ts32"; timeout 900 python tools/sweep.py 60 500 32 0 5000 > gpurun_out/sweep_60_500_ts32.log 2>&1; cat gpurun_out/sweep_60_500_ts32.log
echo "=== sweep caffeine"; timeout 900 python tools/sweep.py 51 195 28 0 1500 > gpurun_out/sweep_caffeine.log 2>&1; cat gpurun_out/sweep_caffeine.log
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
