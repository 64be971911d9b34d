#!/bin/bash
# Round-2 final check (1 GPU, the tree as committed): GPU tests incl. the large fixtures, smoke, the default bench line,
# caffeine and benzene whole jobs with the final kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all_r02.log 2>&1; tail -3 gpurun_out/t_all_r02.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; tail -1 gpurun_out/smoke_r02.log
echo "=== bench N=1 (default flags)"; timeout 1500 python bench.py > gpurun_out/bench_r02_n1_final.json 2> gpurun_out/bench_r02_n1_final.err; cut -c1-200 gpurun_out/bench_r02_n1_final.json; tail -c 300 gpurun_out/bench_r02_n1_final.err
echo "=== caffeine / benzene / uracil whole jobs"; 
timeout 600 python tools/exec_sweep.py 51 195 28 0 -1,0 > gpurun_out/final_caffeine.jsonl 2>&1
timeout 600 python tools/exec_sweep.py 21 93 40 0 0 > gpurun_out/final_benzene.jsonl 2>&1
timeout 600 python tools/exec_sweep.py 29 103 40 0 -1 > gpurun_out/final_uracil.jsonl 2>&1
cat gpurun_out/final_*.jsonl | cut -c1-60,330-760
