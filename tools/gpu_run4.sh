#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_n60v500 -f python tools/run_one.py 60 500 32 0 1,1,2 2 > gpurun_out/ncu_n60.log 2>&1; tail -3 gpurun_out/ncu_n60.log
