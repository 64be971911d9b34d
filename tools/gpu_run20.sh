#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export SWEEP_CONFIGS="1,1,1"
for v in u1 u2 u4; do
  export CCSDT_B200_LIB=$PWD/tools/_ab/lib_$v.so
  echo "=== $v benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 27 2>&1 | cut -c80-400
  echo "=== $v 60/500"; timeout 600 python tools/sweep.py 60 500 32 0 5000 2>&1 | cut -c80-400
  echo "=== $v caffeine"; timeout 600 python tools/sweep.py 51 195 28 0 1500 2>&1 | cut -c80-400
done > gpurun_out/ab_unroll.log 2>&1
cat gpurun_out/ab_unroll.log
