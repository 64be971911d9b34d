"""Execution-tiling sweep (ccsdt_options.exec_tilesize): the same problem on the caller's tiles and on re-cut execution
tiles.  Procedural tensors generated on the device; a strided sample of the EXECUTION task list, or the whole job.
  usage: python tools/exec_sweep.py noa nva ts ntasks exec_ts[,exec_ts...] [symmetry]
  ntasks = 0: whole job.  Prints one JSON line per execution tile size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import driver as drv, synthetic as syn  # noqa: E402

no, nv, ts, ntasks = (int(x) for x in sys.argv[1:5])
targets = [int(x) for x in sys.argv[5].split(",")]
symmetry = int(sys.argv[6]) if len(sys.argv) > 6 else 1
sp = drv.setup_mo_space(no, no, nv, nv, ts)
evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
ref_ops = float(drv.count_ops(sp, True))
for target in targets:
    ctx = drv.Context(0)
    ctx.set_options(exec_tilesize=target, symmetry=symmetry)
    ctx.set_space(sp, evl, True)
    ctx.set_synthetic(1234)
    ex = ctx.exec_space()
    n = ctx.num_tasks()
    ids = np.arange(n, dtype=np.int64) if ntasks <= 0 or ntasks >= n else (np.arange(ntasks, dtype=np.int64) * n) // ntasks
    ctx.run_tasks(ids[:2])                        # warm-up: pools, box lists
    t0 = time.perf_counter()
    e1, e2, st, _ = ctx.run_tasks(ids)
    dt = time.perf_counter() - t0
    ctx.close()
    whole = len(ids) == n
    print(json.dumps({
        "problem": f"({no},{nv}) ts{ts}", "exec_tilesize": target, "symmetry": symmetry,
        "exec_tiles_occ_alpha": [int(x) for x in ex.k_range[:ex.noa]], "exec_tiles_virt_alpha": [int(x) for x in ex.k_range[ex.noab:ex.noab + ex.nva]],
        "kernel_tasks": int(len(ids)), "of": int(n), "wall_s": dt, "kernel_s": st["seconds_kernel"], "staging_s": st["seconds_staging"],
        "launches": st["kernel_launches"],
        "counted_tflops_reference_tiling": ref_ops / dt / 1e12 if whole else None,
        "counted_tflops_exec_tiling": st["counted_flops"] / dt / 1e12,
        "kernel_tflops_evaluated": st["evaluated_flops"] / st["seconds_kernel"] / 1e12,
        "kernel_tflops_executed": st["executed_flops"] / st["seconds_kernel"] / 1e12,
        "executed_over_evaluated": st["executed_flops"] / st["evaluated_flops"],
        "evaluated_over_counted": st["evaluated_flops"] / st["counted_flops"],
        "E[T]": e1, "E(T)": e2}), flush=True)
