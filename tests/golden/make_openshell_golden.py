"""Generates tests/golden/ch2_triplet_321g.npz: UHF / spin-orbital CCSD amplitudes and integrals of triplet methylene
(geometry of the reference's inputs/ch2.json, 3-21G so that the dense spin-orbital tensors stay small) from tools/provider,
and the energies of the reference's own CPU (T) path on it (is_restricted = false, one tile per spin block: the reference's
k_spin rule needs equal alpha / beta tile counts) -> tests/golden/molecules_ref.json.

  python tests/golden/make_openshell_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tools.provider import provider as pv  # noqa: E402

REF = "/root/reference"

if __name__ == "__main__":
    T, info = pv.solve_uhf(os.path.join(REF, "inputs/ch2.json"), os.path.join(REF, "exachem/basis"), 3, basisset="3-21g")
    info["input"], info["basisset"], info["multiplicity"] = "inputs/ch2.json", "3-21g", 3
    np.savez(os.path.join(HERE, "ch2_triplet_321g.npz"), summary=json.dumps(info), **T)
    from oracle.oracle import Oracle, Reference
    na, nb, n = info["n_occ_alpha"], info["n_occ_beta"], info["nbf"]
    orc = Oracle()
    sp = orc.tiles(na, nb, n - na, n - nb, 40)
    e, trace = Reference().execute(sp, T, False, tilesize=40)
    path = os.path.join(HERE, "molecules_ref.json")
    out = json.load(open(path))
    out["ch2_triplet_321g"] = {"ccsdt_tilesize": 40, "E[T]": float(e[0]), "E(T)": float(e[1]), "kernel_tasks": int(len(trace)),
                               "is_restricted": False,
                               "source": "CCSD_T_Fused_Driver<double>::execute + total_fused_ccsd_t_cpu (oracle/_ref) on the fixture"}
    json.dump(out, open(path, "w"), indent=1)
    print(info, out["ch2_triplet_321g"])
