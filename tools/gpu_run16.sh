#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export SWEEP_DEBUG="0,16,32,48,15,63"
echo "=== sweep 60/500 debug"; timeout 600 python tools/sweep.py 60 500 32 0 > gpurun_out/sweep_n60_dbg.log 2>&1; cat gpurun_out/sweep_n60_dbg.log
echo "=== sweep benzene debug"; timeout 600 python tools/sweep.py 21 93 40 0 > gpurun_out/sweep_benzene_dbg.log 2>&1; cat gpurun_out/sweep_benzene_dbg.log
