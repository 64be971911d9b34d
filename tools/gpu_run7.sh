#!/bin/bash
# WS kernel + brick-ordered dynamic box scheduler: tests, sweeps, ncu captures (benzene task 0, 60/500 ts32 task 5000)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
echo "=== sweep benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 14 27 > gpurun_out/sweep_benzene.log 2>&1; cat gpurun_out/sweep_benzene.log
echo "=== sweep 60/500 ts32"; timeout 900 python tools/sweep.py 60 500 32 0 5000 12000 > gpurun_out/sweep_60_500_ts32.log 2>&1; cat gpurun_out/sweep_60_500_ts32.log
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
echo "=== ncu full 60/500"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01c_n60v500 -f python tools/run_one.py 60 500 32 5000 1,1,1 2 > gpurun_out/ncu_n60.log 2>&1; tail -3 gpurun_out/ncu_n60.log
echo "=== ncu full benzene"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01c_benzene -f python tools/run_one.py 21 93 40 0 1,1,1 2 > gpurun_out/ncu_benz.log 2>&1; tail -3 gpurun_out/ncu_benz.log
ls -la gpurun_out
