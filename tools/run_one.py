"""Runs single kernel tasks of a synthetic problem (for ncu).  usage: run_one.py NOCC NVIR TS TASK [sub a,b,c] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import driver as drv, synthetic as syn
no, nv, ts, task = (int(x) for x in sys.argv[1:5])
sub = tuple(int(x) for x in sys.argv[5].split(",")) if len(sys.argv) > 5 else (1, 1, 2)
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
sp = drv.setup_mo_space(no, no, nv, nv, ts)
ctx = drv.Context(0)
ctx.set_options(sub=sub)
ctx.set_space(sp, syn.Orbitals(no, no, nv, nv).orbital_energies(), True)
ctx.set_synthetic(1234)
for _ in range(reps):
    e1, e2, st, _ = ctx.run(task, task + 1)
    print(e1, e2, st["counted_flops"] / st["seconds_kernel"] / 1e12, "TF/s", st["seconds_kernel"] * 1e3, "ms")
ctx.close()
