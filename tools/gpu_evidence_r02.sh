#!/bin/bash
# Round-2 evidence run (1 GPU): the default bench line, the ncu launch list of the bench command, ncu --set full of the fused
# kernel on one benzene task, one caffeine task on execution tiles of 40 and (60,500) task 5000, the BASELINE configs[4] sweep.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== bench N=1 (default flags)"; timeout 1500 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; cut -c1-300 gpurun_out/bench_r02_n1.json; tail -c 300 gpurun_out/bench_r02_n1.err
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_reference_arm.json
echo "=== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-benzene > gpurun_out/bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu.log
echo "=== ncu full: benzene task 8"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r02_benzene_task8 -f python tools/run_one.py 21 93 40 8 1,1,1 2 > gpurun_out/ncu_benz.log 2>&1; tail -2 gpurun_out/ncu_benz.log
echo "=== ncu full: caffeine, execution tiles of 40"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r02_caffeine_exec40 -f python tools/run_one.py 51 195 28 -1 1,1,1 2 40 > gpurun_out/ncu_caf.log 2>&1; tail -3 gpurun_out/ncu_caf.log
echo "=== ncu full: caffeine, its own tiles (28)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r02_caffeine_ts28 -f python tools/run_one.py 51 195 28 -1 1,1,1 2 0 > gpurun_out/ncu_caf28.log 2>&1; tail -3 gpurun_out/ncu_caf28.log
echo "=== ncu full: (60,500) task 5000"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_t_dmma -s 1 -c 1 -o gpurun_out/prof_r02_n60v500_task5000 -f python tools/run_one.py 60 500 32 5000 1,1,1 2 > gpurun_out/ncu_n60.log 2>&1; tail -2 gpurun_out/ncu_n60.log
echo "=== ncu full: panel build of a (60,500) ts16 task on execution tiles of 40 (blocks cut and merged)"; timeout 600 ncu --set full --clock-control none -k regex:gather_panels -s 1 -c 1 -o gpurun_out/prof_r02_gather_retiled -f python tools/run_one.py 60 500 16 -1 1,1,1 2 40 > gpurun_out/ncu_gather.log 2>&1; tail -2 gpurun_out/ncu_gather.log
echo "=== sweep (BASELINE configs[4])"
for cfg in "20 200 16 0" "20 200 24 0" "40 300 32 24" "80 650 40 12" "100 800 48 6" "100 800 64 2" "60 500 24 24" "60 500 40 24" "60 500 48 12" "60 500 64 4"; do
  timeout 600 python tools/exec_sweep.py $cfg -1 >> gpurun_out/sweep_r02.jsonl 2>> gpurun_out/sweep_r02.err
done
cut -c1-60,330-520 gpurun_out/sweep_r02.jsonl
