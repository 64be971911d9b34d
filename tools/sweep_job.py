"""Whole-job sweep over launch options: kernel and wall time of every kernel task of a synthetic problem.
usage: python tools/sweep_job.py NOCC NVIR TS [max_tasks]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from exachem_b200 import driver as drv, synthetic as syn


def main():
    no, nv, ts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    sp = drv.setup_mo_space(no, no, nv, nv, ts)
    tasks, _, _ = drv.enumerate_tasks(sp, True)
    n = min(len(tasks), int(sys.argv[4])) if len(sys.argv) > 4 else len(tasks)
    ids = (np.arange(n, dtype=np.int64) * len(tasks)) // n
    evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
    configs = [dict(), dict(stagger=0), dict(ctas_per_sm=2), dict(ctas_per_sm=2, stagger=0), dict(sub=(1, 1, 2)),
               dict(sub=(1, 2, 1)), dict(sub=(2, 1, 1)), dict(stages=2), dict(overlap=0), dict(symmetry=0)]
    for cfg in configs:
        ctx = drv.Context(0)
        ctx.set_options(**cfg)
        ctx.set_space(sp, evl, True)
        ctx.set_synthetic(1234)
        ctx.run_tasks(ids)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            e1, e2, st, _ = ctx.run_tasks(ids)
            wall = time.perf_counter() - t0
            if best is None or wall < best[0]:
                best = (wall, st)
        ctx.close()
        wall, st = best
        print(json.dumps({"cfg": {k: list(v) if isinstance(v, tuple) else v for k, v in cfg.items()},
                          "wall_ms": round(wall * 1e3, 2), "kernel_ms": round(st["seconds_kernel"] * 1e3, 2),
                          "counted_TF/s_wall": round(st["counted_flops"] / wall / 1e12, 2),
                          "evaluated_TF/s_kernel": round(st["evaluated_flops"] / st["seconds_kernel"] / 1e12, 2),
                          "E(T)": e2}), flush=True)


main()
