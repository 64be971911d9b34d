#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export SWEEP_DEBUG="0,1,3,7"
echo "=== sweep benzene debug"; timeout 600 python tools/sweep.py 21 93 40 0 27 > gpurun_out/sweep_benzene_dbg.log 2>&1; cat gpurun_out/sweep_benzene_dbg.log
