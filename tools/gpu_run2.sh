#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== sweep benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 14 27 > gpurun_out/sweep_benzene.log 2>&1; cat gpurun_out/sweep_benzene.log
echo "=== sweep 60/500 ts32"; timeout 900 python tools/sweep.py 60 500 32 0 5000 12000 > gpurun_out/sweep_60_500_ts32.log 2>&1; cat gpurun_out/sweep_60_500_ts32.log
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 30 -c 3 -o gpurun_out/prof_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1; tail -2 gpurun_out/bench_under_ncu_full.log | cut -c1-200
ls -la gpurun_out
