#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
echo "=== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
echo "=== bench synth 1gpu"; timeout 900 python bench.py --workload synth --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_synth1.log 2>&1; tail -1 gpurun_out/bench_synth1.log | cut -c1-1500
