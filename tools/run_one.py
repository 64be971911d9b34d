"""Runs single kernel tasks of a synthetic problem (for ncu).
usage: run_one.py NOCC NVIR TS TASK [sub a,b,c] [reps] [exec_tilesize]
TASK < 0: the first task of the (execution) task list whose six tiles are all different and as large as tiles get."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import driver as drv, synthetic as syn
no, nv, ts, task = (int(x) for x in sys.argv[1:5])
sub = tuple(int(x) for x in sys.argv[5].split(",")) if len(sys.argv) > 5 else (1, 1, 1)
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
exec_ts = int(sys.argv[7]) if len(sys.argv) > 7 else 0
sp = drv.setup_mo_space(no, no, nv, nv, ts)
ctx = drv.Context(0)
ctx.set_options(sub=sub, exec_tilesize=exec_ts)
ctx.set_space(sp, syn.Orbitals(no, no, nv, nv).orbital_energies(), True)
ctx.set_synthetic(1234)
ex = ctx.exec_space()
if task < 0:
    tasks, _, _ = drv.enumerate_tasks(ex, True)
    vol = np.array([np.prod(ex.k_range[t[:6]]) if len(set(t[:6])) == 6 else 0 for t in tasks], dtype=float)
    task = int(np.argmax(vol))
    print("task", task, tasks[task][:6], "extents", ex.k_range[tasks[task][:6]])
for _ in range(reps):
    e1, e2, st, _ = ctx.run_tasks([task])
    print(e1, e2, st["evaluated_flops"] / st["seconds_kernel"] / 1e12, "TF/s evaluated", st["executed_flops"] / st["seconds_kernel"] / 1e12,
          "TF/s executed", st["seconds_kernel"] * 1e3, "ms")
ctx.close()
