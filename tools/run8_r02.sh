set -x
nvidia-smi topo -m | head -12; free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02_n8_synth.json 2> gpurun_out/bench_r02_n8_synth.err
tail -c 400 gpurun_out/bench_r02_n8_synth.err
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r02_n4_synth.json 2> gpurun_out/bench_r02_n4_synth.err
tail -c 400 gpurun_out/bench_r02_n4_synth.err
$TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload caffeine --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_r02_n8_caffeine_dynamic.json 2> gpurun_out/bench_r02_n8_caffeine_dynamic.err
tail -c 400 gpurun_out/bench_r02_n8_caffeine_dynamic.err
$TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --workload caffeine --static --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_r02_n8_caffeine_static.json 2> gpurun_out/bench_r02_n8_caffeine_static.err
tail -c 400 gpurun_out/bench_r02_n8_caffeine_static.err
$TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --workload gc --tasks 320 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_r02_n8_gc.json 2> gpurun_out/bench_r02_n8_gc.err
tail -c 400 gpurun_out/bench_r02_n8_gc.err
python -m pytest tests/test_multirank_cpp.py -q -m gpu 2>&1 | tail -3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r02_n*_*.json")):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, d["n_gpus"], round(d["value"],2), round(d["ms_per_step"],1), round(d["roofline"]["frac"],3), round(d["symmetry"]["value_symmetry_off"],2), {k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","h2d_bytes_per_step","peer_bytes_per_step")})
    except Exception as e:
        print(f, "ERR", e)
PY
