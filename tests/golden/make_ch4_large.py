"""Generates tests/golden/_large/ch4_def2tzvp_uhf.npz (~90 MB, git-ignored; travels to the GPU box with the snapshot):
the UHF / spin-orbital CCSD amplitudes and integrals of the reference's open-shell CI case inputs/ci/ch4.json (triplet
methane, def2-TZVP with f functions, 6 alpha / 4 beta electrons, ccsdt_tilesize 28) from tools/provider, as dense
spin-orbital tensors, plus the oracle's energies on them -> tests/golden/molecules_ref.json["ch4_def2tzvp_uhf"].
The reference's own CI golden for this case (ci/reference_output/ch4.def2-tzvp.ccsd_t.json) is what the GPU test checks.

  python tests/golden/make_ch4_large.py
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
    T, info = pv.solve_uhf(os.path.join(REF, "inputs/ci/ch4.json"), os.path.join(REF, "exachem/basis"), 3, verbose=False)
    info["input"], info["multiplicity"], info["ccsdt_tilesize"] = "inputs/ci/ch4.json", 3, 28
    os.makedirs(os.path.join(HERE, "_large"), exist_ok=True)
    np.savez(os.path.join(HERE, "_large", "ch4_def2tzvp_uhf.npz"), summary=json.dumps(info), **T)
    from oracle.oracle import Oracle
    orc = Oracle()
    na, nb, n = info["n_occ_alpha"], info["n_occ_beta"], info["nbf"]
    e1, e2 = orc.run(orc.tiles(na, nb, n - na, n - nb, 28), T, False)
    path = os.path.join(HERE, "molecules_ref.json")
    out = json.load(open(path))
    out["ch4_def2tzvp_uhf"] = {"ccsdt_tilesize": 28, "E[T]": float(e1), "E(T)": float(e2), "is_restricted": False,
                               "n_occ_alpha": na, "n_occ_beta": nb, "nbf": n,
                               "source": "oracle/ccsdt_oracle.c (pinned bit for bit to the reference CPU path) on the fixture"}
    json.dump(out, open(path, "w"), indent=1)
    print(info, out["ch4_def2tzvp_uhf"])
