"""Runs the first kernel tasks of a synthetic problem through the reference's own GPU path (oracle/_ref, K1 DMMA kernel)
-- for ncu.  TEST/MEASUREMENT TOOL.  usage: run_one_reference.py NOCC NVIR TS NTASKS"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import synthetic as syn
from oracle.oracle import Oracle, ReferenceGPU

no, nv, ts, n = (int(x) for x in sys.argv[1:5])
orc, ref = Oracle(), ReferenceGPU("tc")
osp = orc.tiles(no, no, nv, nv, ts)
out, _ = ref.execute_synth(osp, syn.Orbitals(no, no, nv, nv).orbital_energies(), [no, no, nv, nv], 1234, True, tilesize=ts, task_limit=n)
for r in ref.last_trace:
    print("extents", r[:6].astype(int).tolist(), "blocks", int(r[6]), "kernel ms %.2f" % r[7])
