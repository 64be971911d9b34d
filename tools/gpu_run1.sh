#!/bin/bash
# first GPU contact: probes, parity tests in separate processes (a trapped kernel kills its CUDA context), quick bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run t_probes python -m pytest tests/test_gpu_parity.py -q -s -k "dmma_fragment or swizzle or synthetic_generator or fp64_peak"
run t_simple python -m pytest tests/test_gpu_parity.py -q -k "fixture and simple"
run t_dmma_fixture python -m pytest tests/test_gpu_parity.py -q -k "fixture and dmma"
run t_dmma_rest python -m pytest tests/test_gpu_parity.py -q -k "not fixture and not dmma_fragment and not swizzle and not synthetic_generator and not fp64_peak"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 3 --warmup 3 --no-cpu-baseline
