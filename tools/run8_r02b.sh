set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02b_n8_synth.json 2> gpurun_out/bench_r02b_n8_synth.err
tail -c 300 gpurun_out/bench_r02b_n8_synth.err
$TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --workload caffeine --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_r02b_n8_caffeine.json 2> gpurun_out/bench_r02b_n8_caffeine.err
tail -c 300 gpurun_out/bench_r02b_n8_caffeine.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r02b_n8_*.json")):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, d["n_gpus"], round(d["value"],2), round(d["ms_per_step"],1), round(d["roofline"]["frac"],3), round(d["symmetry"]["value_symmetry_off"],2), {k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","h2d_bytes_per_step","peer_bytes_per_step")})
    except Exception as e:
        print(f, "ERR", e)
PY
