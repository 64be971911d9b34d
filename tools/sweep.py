"""Kernel-level sweep over CTA box shapes / stages on sampled tasks of a synthetic problem.
usage: python tools/sweep.py NOCC NVIR TS [task indices...]   (prints TFLOP/s per config)"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from exachem_b200 import driver as drv, synthetic as syn

def main():
    no, nv, ts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    sp = drv.setup_mo_space(no, no, nv, nv, ts)
    tasks, fac, _ = drv.enumerate_tasks(sp, True)
    idx = [int(x) for x in sys.argv[4:]] or [0, len(tasks) // 3, (2 * len(tasks)) // 3, len(tasks) - 1]
    evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
    configs = [dict(sub=(1, 1, 1)), dict(sub=(1, 1, 1), stagger=0), dict(sub=(1, 1, 2)), dict(sub=(1, 1, 1), ctas_per_sm=2),
               dict(sub=(1, 1, 1), ctas_per_sm=2, stagger=0)]
    if os.environ.get("SWEEP_CONFIGS"):
        configs = [dict(sub=tuple(int(x) for x in c.split(","))) for c in os.environ["SWEEP_CONFIGS"].split(";")]
    for cfg in configs:
        ctx = drv.Context(0)
        ctx.set_options(**cfg)
        ctx.set_space(sp, evl, True)
        ctx.set_synthetic(1234)
        out = []
        for i in idx:
            ctx.run(i, i + 1)  # warm
            e1, e2, st, _ = ctx.run(i, i + 1)
            out.append((i, tasks[i][:6].tolist(), sp.k_range[tasks[i][:6]].tolist(), round(st["counted_flops"] / st["seconds_kernel"] / 1e12, 2),
                        round(st["seconds_kernel"] * 1e3, 2), round(st["seconds_staging"] * 1e3, 2)))
        ctx.close()
        print(json.dumps({"cfg": {k: list(v) if isinstance(v, tuple) else v for k, v in cfg.items()}, "tasks(idx,tiles,ext,TF/s,kernel_ms,staging_ms)": out}))

main()
