"""Generates tests/golden/{h2o_ccpvdz,butanol2_sto3g}.npz: converged RHF/CCSD amplitudes and MO integrals of two of
the reference's own inputs, produced in THIS container by tools/provider (geometry and basis-set tables are read
from /root/reference at generation time; nothing of it is copied into the repo except the numbers computed here).

  python tests/golden/make_molecule_golden.py

Checks printed at generation time (and asserted in tests/test_molecules.py from the stored summaries):
  H2O cc-pVDZ (inputs/h2o.json)          SCF  -75.82509922164868  CCSD corr -0.25498209984722586
      (ci/reference_output/h2o_eom.cc-pvdz.eom_ccsd.json:139,265-268; CCSD threshold there is 1e-6)
  butanol2 STO-3G (inputs/ci/butanol2_pt.json)  SCF -229.2941781065417, CCSD corr -0.32235871392598026,
      [T] -0.002394401847429249, (T) -0.002256149503764586 (ci/reference_output/butanol2_pt.sto-3g.ccsd_t.json)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tools.provider import provider as pv  # noqa: E402

REF = "/root/reference"
CASES = {"h2o_ccpvdz": ("inputs/h2o.json", 28), "butanol2_sto3g": ("inputs/ci/butanol2_pt.json", 40)}

if __name__ == "__main__":
    for name, (inp, ts) in CASES.items():
        fx, info = pv.solve(os.path.join(REF, inp), os.path.join(REF, "exachem/basis"))
        info["input"], info["ccsdt_tilesize"] = inp, ts
        np.savez(os.path.join(HERE, name + ".npz"), summary=json.dumps(info), **fx)
        print(name, "written")
    # the reference's OWN CPU (T) path (oracle/_ref, compiled unmodified) on these amplitudes -> molecules_ref.json
    from oracle.oracle import Oracle, Reference
    path = os.path.join(HERE, "molecules_ref.json")
    ref, orc = Reference(), Oracle()
    out = json.load(open(path)) if os.path.exists(path) else {}   # keeps entries made elsewhere (tools/benzene_real.py)
    for name, (inp, ts) in CASES.items():
        fx = np.load(os.path.join(HERE, name + ".npz"))
        T = pv.spin_orbital_tensors(fx)
        no, nv = int(fx["nocc"]), len(fx["eps"]) - int(fx["nocc"])
        sp = orc.tiles(no, no, nv, nv, ts)
        e, trace = ref.execute(sp, T, True, tilesize=ts)
        out[name] = {"ccsdt_tilesize": ts, "E[T]": float(e[0]), "E(T)": float(e[1]), "kernel_tasks": int(len(trace)),
                     "source": "CCSD_T_Fused_Driver<double>::execute + total_fused_ccsd_t_cpu (oracle/_ref) on the fixture"}
        print(name, out[name])
    json.dump(out, open(path, "w"), indent=1)
