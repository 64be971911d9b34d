#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
echo "=== sweep benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 14 27 > gpurun_out/sweep_benzene.log 2>&1; cat gpurun_out/sweep_benzene.log
echo "=== sweep 60/500 ts32"; timeout 900 python tools/sweep.py 60 500 32 0 5000 > gpurun_out/sweep_60_500_ts32.log 2>&1; cat gpurun_out/sweep_60_500_ts32.log
