#!/bin/bash
# Round-1 evidence run (1 GPU): tests, bench lines, ncu launch list of the bench command, ncu --set full of the top kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
echo "=== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
echo "=== bench synth 1gpu"; timeout 900 python bench.py --workload synth --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_synth1.log 2>&1; tail -1 gpurun_out/bench_synth1.log | cut -c1-300
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 168 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120
echo "=== ncu full 60/500"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_n60v500 -f python tools/run_one.py 60 500 32 5000 1,1,1 2 > gpurun_out/ncu_n60.log 2>&1; tail -2 gpurun_out/ncu_n60.log
echo "=== ncu full benzene"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_benzene -f python tools/run_one.py 21 93 40 0 1,1,1 2 > gpurun_out/ncu_benz.log 2>&1; tail -2 gpurun_out/ncu_benz.log
echo "=== ncu full gather"; timeout 600 ncu --set full --clock-control none -k regex:gather_panels -s 1 -c 1 -o gpurun_out/prof_r01_gather -f python tools/run_one.py 60 500 32 5000 1,1,1 2 > gpurun_out/ncu_gather.log 2>&1; tail -1 gpurun_out/ncu_gather.log
