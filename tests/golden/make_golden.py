"""Generates tests/golden/ref_small.json from the REFERENCE's own code (oracle/_ref, i.e.
/root/reference/exachem/cc/ccsd_t compiled unmodified against oracle/shim).  Run in the build
container (needs /root/reference):  python tests/golden/make_golden.py

Each case: synthetic tensors (exachem_b200.synthetic, seed recorded), the reference's energies
(17 significant digits), its kernel-task trace (h1b,h2b,h3b,p4b,p5b,p6b,outer id,factor), its
total_num_ops, and the exec tables of every task from its *_info_only functions.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from exachem_b200 import synthetic as syn  # noqa: E402
from oracle.oracle import Oracle, Reference  # noqa: E402

CASES = [  # name, noa, nob, nva, nvb, tilesize, restricted, seed
    ("o4v6_ts3", 4, 4, 6, 6, 3, True, 1234),
    ("o4v6_ts4", 4, 4, 6, 6, 4, True, 1234),
    ("o4v6_whole", 4, 4, 6, 6, 6, True, 1234),
    ("o3v9_ts4", 3, 3, 9, 9, 4, True, 7),
    ("o5v11_ts8", 5, 5, 11, 11, 8, True, 99),
    ("uhf_o3o2_v5v6_ts3", 3, 2, 5, 6, 3, False, 5),
    ("uhf_o4v6_ts4", 4, 4, 6, 6, 4, False, 1234),
    # the shape of BASELINE configs[0]: H2O cc-pVDZ (5 occupied, 19 virtual orbitals per spin) at the
    # ccsdt_tilesize of inputs/h2o.json (28): tiles [5][5][19][19], two kernel tasks
    ("h2o_shape_ts28", 5, 5, 19, 19, 28, True, 2024),
    ("h2o_shape_ts7", 5, 5, 19, 19, 7, True, 2024),
]


def main():
    orc, ref = Oracle(), Reference()
    out = {}
    for name, noa, nob, nva, nvb, ts, restricted, seed in CASES:
        sp = orc.tiles(noa, nob, nva, nvb, ts)
        if not restricted and (sp.noa != sp.nob or sp.nva != sp.nvb):
            # the reference labels spins half/half (ccsd_t.cpp:245-249); keep cases it handles
            continue
        T = syn.dense_all(syn.Orbitals(noa, nob, nva, nvb), seed)
        res, trace = ref.execute(sp, T, restricted, tilesize=ts)
        execs = []
        for t in trace:
            s1, d1, d2, cnt = ref.task_info(sp, restricted, t[:6].astype(int))
            execs.append({"s1": s1.tolist(), "d1": d1.tolist(), "d2": d2.tolist(), "n": cnt.tolist()})
        out[name] = {
            "noa": noa, "nob": nob, "nva": nva, "nvb": nvb, "tilesize": ts, "restricted": restricted,
            "seed": seed, "k_range": sp.k_range.tolist(), "k_spin": sp.k_spin.tolist(),
            "energy1": repr(float(res[0])), "energy2": repr(float(res[1])),
            "tasks": trace[:, :7].astype(int).tolist(), "factors": [repr(float(x)) for x in trace[:, 7]],
            "running_e1": [repr(float(x)) for x in trace[:, 8]],
            "running_e2": [repr(float(x)) for x in trace[:, 9]],
            "total_num_ops": ref.count_ops(sp, restricted), "exec": execs,
        }
        print(name, res[:2], len(trace))
    with open(os.path.join(os.path.dirname(__file__), "ref_small.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
