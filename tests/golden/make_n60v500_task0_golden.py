"""Generates tests/golden/n60v500_ts32_task0.json: the energy of kernel task 0 (tiles all 32: the 32^6 task, 36-fold
symmetric) of the synthetic (nocc, nvir) = (60, 500) problem at ccsdt_tilesize 32 -- BASELINE.json configs[4], the
north_star target size -- from the REFERENCE's own CPU path (oracle/_ref: CCSD_T_Fused_Driver<double>::execute +
total_fused_ccsd_t_cpu, unmodified) on the procedural tensors of seed 1234.  32^6 = 1.07e9 elements is the largest tile
the reference CPU kernel can index (int t3 index, ccsd_t_all_fused_cpu.hpp:165); 1.08e13 flops, about two hours here.

  OMP_NUM_THREADS=6 python tests/golden/make_n60v500_task0_golden.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from exachem_b200 import synthetic as syn  # noqa: E402
from oracle.oracle import Oracle, Reference  # noqa: E402

if __name__ == "__main__":
    no, nv, ts, seed = 60, 500, 32, 1234
    orc, ref = Oracle(), Reference()
    osp = orc.tiles(no, no, nv, nv, ts)
    evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
    t0 = time.time()
    out, n = ref.execute_synth(osp, evl, [no, no, nv, nv], seed, True, tilesize=ts, task_limit=1)
    rec = {"problem": "(60,500) ts32, seed 1234, is_restricted", "task": [0, 0, 0, 4, 4, 4], "factor": 2.0 / 36.0,
           "E[T]": float(out[0]), "E(T)": float(out[1]), "seconds": time.time() - t0, "threads": ref.num_threads(),
           "source": "oracle/_ref (the reference's execute + CPU kernel, unmodified), task_limit = 1"}
    json.dump(rec, open(os.path.join(HERE, "n60v500_ts32_task0.json"), "w"), indent=1)
    print(rec)
