#!/bin/bash
# Round-1 final evidence run (1 GPU): tests, smoke, bench lines, ncu launch list of the bench command, ncu --set full of the
# top kernel on the N=1 and N>1 workloads (symmetry on), the on-box GPU comparator, benzene on real amplitudes.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench_r01_n1_benzene.json; cut -c1-200 gpurun_out/bench_r01_n1_benzene.json
echo "=== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log > gpurun_out/bench_r01_reference_arm.json; cut -c1-200 gpurun_out/bench_r01_reference_arm.json
echo "=== bench synth 1gpu"; timeout 900 python bench.py --workload synth --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_synth1.log 2>&1; tail -1 gpurun_out/bench_synth1.log > gpurun_out/bench_r01_n1_synth60x500.json; cut -c1-200 gpurun_out/bench_r01_n1_synth60x500.json
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120
echo "=== ncu full benzene task0"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_benzene -f python tools/run_one.py 21 93 40 0 1,1,1 2 > gpurun_out/ncu_benz.log 2>&1; tail -2 gpurun_out/ncu_benz.log
echo "=== ncu full 60/500 task0"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_n60v500_t0 -f python tools/run_one.py 60 500 32 0 1,1,1 2 > gpurun_out/ncu_n60a.log 2>&1; tail -2 gpurun_out/ncu_n60a.log
echo "=== ncu full 60/500 task5000"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r01_n60v500 -f python tools/run_one.py 60 500 32 5000 1,1,1 2 > gpurun_out/ncu_n60.log 2>&1; tail -2 gpurun_out/ncu_n60.log
echo "=== comparator"; timeout 900 python tools/gpu_comparator.py --out gpurun_out/comparator_r01.json > gpurun_out/comparator.log 2>&1; tail -3 gpurun_out/comparator.log
if [ -f tests/golden/_large/benzene_ccpvdz.npz ]; then
  echo "=== benzene real"; timeout 900 python tools/benzene_real.py --gpu --out gpurun_out/benzene_real_r01.json > gpurun_out/benzene_real.log 2>&1; tail -4 gpurun_out/benzene_real.log
fi
