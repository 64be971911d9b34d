#!/bin/bash
# re-entry run: full GPU test suite, peaks/mainloop probe, kernel sweeps, bench, ncu launch list + one full capture
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "=== probe mainloop"; timeout 600 python tools/probe_mainloop.py > gpurun_out/probe_mainloop.log 2>&1; cat gpurun_out/probe_mainloop.log
echo "=== sweep benzene"; timeout 600 python tools/sweep.py 21 93 40 0 5 14 27 > gpurun_out/sweep_benzene.log 2>&1; cat gpurun_out/sweep_benzene.log
echo "=== sweep 60/500 ts32"; timeout 900 python tools/sweep.py 60 500 32 0 5000 12000 > gpurun_out/sweep_60_500_ts32.log 2>&1; cat gpurun_out/sweep_60_500_ts32.log
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
echo "=== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_n60v500 -f python tools/run_one.py 60 500 32 5000 1,1,2 2 > gpurun_out/ncu_n60.log 2>&1; tail -3 gpurun_out/ncu_n60.log
ls -la gpurun_out
