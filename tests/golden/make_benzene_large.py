"""Generates tests/golden/_large/benzene_ccpvdz.npz (~200 MB, git-ignored; it travels to the GPU box with the
snapshot like the built libraries): converged RHF/CCSD amplitudes and MO integrals of the reference's
inputs/benzene.json (BASELINE.json configs[1]) from tools/provider, in the same compact closed-shell form as
the small fixtures.  About one hour of CPU (spin-orbital CCSD with 228 spin orbitals, ~35 GB of RAM).

  OMP_NUM_THREADS=6 python tests/golden/make_benzene_large.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tools.provider import provider as pv  # noqa: E402

REF = "/root/reference"

if __name__ == "__main__":
    t0 = time.time()
    fx, info = pv.solve(os.path.join(REF, "inputs/benzene.json"), os.path.join(REF, "exachem/basis"), verbose=True)
    info["input"], info["ccsdt_tilesize"], info["seconds"] = "inputs/benzene.json", 40, time.time() - t0
    os.makedirs(os.path.join(HERE, "_large"), exist_ok=True)
    np.savez(os.path.join(HERE, "_large", "benzene_ccpvdz.npz"), summary=json.dumps(info), **fx)
    json.dump(info, open(os.path.join(HERE, "benzene_ccpvdz_summary.json"), "w"), indent=1)
    print("written", info)
