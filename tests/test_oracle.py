"""Pins the oracle (oracle/ccsdt_oracle.c) -- CPU only.

1. integer known-answer tests: the six `total_num_ops` values in the reference's CI goldens
   (/root/reference/ci/reference_output/*.ccsd_t.json, cited per case);
2. committed fixtures generated from the reference's own code (tests/golden/ref_small.json, made by
   tests/golden/make_golden.py): task lists, factors, exec tables bit-exact, energies bit-exact;
3. live comparison with oracle/_ref (the unmodified reference compiled against the TAMM shim) when
   that library is present;
4. an independent closed-form einsum statement and tiling invariance.
"""
import json
import os

import numpy as np
import pytest

from exachem_b200 import synthetic as syn
from oracle.oracle import closed_form_energy

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_small.json")))

# (name, n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, ccsdt_tilesize, restricted, total_num_ops, source)
OPS_KAT = [
    ("butanol2", 21, 21, 14, 14, 40, True, 30952040112, "butanol2_pt.sto-3g.ccsd_t.json:260-263"),
    ("ch4_uhf", 6, 4, 49, 51, 28, False, 37432196256, "ch4.def2-tzvp.ccsd_t.json:482-485"),
    ("seh2_fc4", 9, 9, 36, 36, 40, True, 52991044992, "seh2.cc-pvdz.ccsd_t.json:261-264"),
    ("uracil", 29, 29, 103, 103, 40, True, 62789556886348, "uracil.cc-pvdz.ccsd_t.json:381-384"),
    ("ubiquitin_631g", 146, 146, 278, 278, 40, True, 2.0292727538577997e17,
     "ubiquitin_dgrtl/ubiquitin_dgrtl.6-31g.ccsd_t.json:285-288"),
    ("ubiquitin_ccpvdz", 146, 146, 591, 591, 40, True, 3.088106365505672e18,
     "ubiquitin_dgrtl/ubiquitin_dgrtl.cc-pvdz.ccsd_t.json:277-280"),
]


@pytest.mark.parametrize("case", OPS_KAT, ids=[c[0] for c in OPS_KAT])
def test_total_num_ops_known_answers(orc, case):
    _, oa, ob, va, vb, ts, restricted, gold, _src = case
    sp = orc.tiles(oa, ob, va, vb, ts)
    ops = orc.count_ops(sp, restricted)
    assert ops == gold if isinstance(gold, int) else float(ops) == gold


def _space_and_tensors(orc, g):
    sp = orc.tiles(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    return sp, T


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_reference_fixture(orc, name):
    g = GOLD[name]
    sp, T = _space_and_tensors(orc, g)
    assert sp.k_range.tolist() == g["k_range"] and sp.k_spin.tolist() == g["k_spin"]
    tasks, fac, _ = orc.enumerate(sp, g["restricted"])
    assert tasks.tolist() == g["tasks"]                      # enumeration bit-exact, in order
    assert [repr(float(f)) for f in fac] == g["factors"]
    assert orc.count_ops(sp, g["restricted"]) == g["total_num_ops"]
    for t, ex in zip(tasks, g["exec"]):
        s1, d1, d2, cnt = orc.task_exec(sp, g["restricted"], t)
        assert s1.tolist() == ex["s1"] and d1.tolist() == ex["d1"] and d2.tolist() == ex["d2"]
        # the reference's *_info_only report the number of enabled TERMS (exec entries >= 0) ...
        assert [int((s1 >= 0).sum()), int((d1 >= 0).sum()), int((d2 >= 0).sum())] == ex["n"]
        # ... while the oracle also returns the number of staged slots (= max slot id + 1)
        assert cnt.tolist() == [int(s1.max()) + 1, int(d1.max()) + 1, int(d2.max()) + 1]
    e1, e2, per_task = orc.run(sp, T, g["restricted"], per_task=True)
    # same operation order as the reference CPU kernel => bit-identical
    assert repr(float(e1)) == g["energy1"] and repr(float(e2)) == g["energy2"]
    run1 = np.cumsum(per_task[:, 0])
    assert np.allclose(run1, [float(x) for x in g["running_e1"]], rtol=0, atol=1e-13)


@pytest.mark.parametrize("cfg", [(4, 4, 6, 6, 3, True, 11), (2, 2, 7, 7, 2, True, 3), (3, 3, 5, 5, 2, False, 8),
                                 (6, 6, 7, 7, 5, True, 21)])
def test_oracle_matches_live_reference(orc, ref, cfg):
    oa, ob, va, vb, ts, restricted, seed = cfg
    sp = orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
    out, trace = ref.execute(sp, T, restricted, tilesize=ts)
    tasks, fac, n_outer = orc.enumerate(sp, restricted)
    assert np.array_equal(trace[:, :7].astype(np.int64), tasks)
    assert np.array_equal(trace[:, 7], fac)
    e1, e2 = orc.run(sp, T, restricted)
    assert e1 == out[0] and e2 == out[1]
    assert orc.count_ops(sp, restricted) == ref.count_ops(sp, restricted)
    for t in tasks:
        a, b = orc.task_exec(sp, restricted, t), ref.task_info(sp, restricted, t)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)


def test_tiling_invariance_and_closed_form(orc):
    oa = ob = 4
    va = vb = 6
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 1234)
    ref_e = closed_form_energy(oa, ob, va, vb, T, True)
    for ts in (2, 3, 4, 6):
        e = orc.run(orc.tiles(oa, ob, va, vb, ts), T, True)
        assert abs(e[0] - ref_e[0]) < 1e-13 and abs(e[1] - ref_e[1]) < 1e-13


def test_unrestricted_closed_form(orc):
    T = syn.dense_all(syn.Orbitals(3, 3, 5, 5), 42)
    ref_e = closed_form_energy(3, 3, 5, 5, T, False)
    e = orc.run(orc.tiles(3, 3, 5, 5, 2), T, False)
    assert abs(e[0] - ref_e[0]) < 1e-13 and abs(e[1] - ref_e[1]) < 1e-13


def test_enumeration_counts_of_named_configs(orc):
    """outer / kernel task counts of the BASELINE configs."""
    for (oa, va, ts, outer, kernel) in [(5, 19, 28, 12, 2), (21, 93, 40, 168, 28), (60, 500, 32, 59840, 16320),
                                        (60, 500, 64, 2448, 408)]:
        sp = orc.tiles(oa, oa, va, va, ts)
        tasks, _, n_outer = orc.enumerate(sp, True)
        assert (n_outer, len(tasks)) == (outer, kernel)
