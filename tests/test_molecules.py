"""(T) on REAL converged amplitudes of two of the reference's own inputs (BASELINE.json configs[0] and a CI case).

Fixtures (tests/golden/{h2o_ccpvdz,butanol2_sto3g}.npz, made by tests/golden/make_molecule_golden.py with
tools/provider: McMurchie-Davidson integrals -> RHF -> spin-orbital CCSD) are pinned to the reference's CI
goldens, hard-coded below with their file:line:
  * ci/reference_output/h2o_eom.cc-pvdz.eom_ccsd.json:139,265-268   (same geometry/basis as inputs/h2o.json)
  * ci/reference_output/butanol2_pt.sto-3g.ccsd_t.json              (SCF, CCSD, [T] and (T) corrections)
The butanol2 golden is the only PUBLISHED (T) energy of the reference that can be reproduced offline (STO-3G needs
no libint); the reference ran it with CD diagtol 1e-5 and a CCSD threshold of 1e-6, which is why agreement is
stated at 5e-9 Eh on the (T) corrections and 1e-6 Eh on the CCSD correlation energy.
tests/golden/molecules_ref.json holds the energies of the reference's own CPU (T) path (oracle/_ref) on the same
fixtures: the 1e-9 Eh bar of BASELINE.json applies against those.
"""
import json
import os

import numpy as np
import pytest

from tools.provider import provider as pv

HERE = os.path.dirname(os.path.abspath(__file__))
REFE = json.load(open(os.path.join(HERE, "golden", "molecules_ref.json")))
ATOL = 1e-9

H2O_GOLD = {"scf": -75.82509922164868, "ccsd_corr": -0.25498209984722586}
BUTANOL_GOLD = {"scf": -229.2941781065417, "ccsd_corr": -0.32235871392598026, "e_nuc": 193.74693406240297,
                "[T]": -0.002394401847429249, "(T)": -0.002256149503764586, "total_num_ops": 30952040112}


def fixture(name):
    fx = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return fx, json.loads(str(fx["summary"]))


def test_fixture_summaries_match_the_reference_ci_goldens():
    _, h = fixture("h2o_ccpvdz")
    assert abs(h["e_scf"] - H2O_GOLD["scf"]) < 1e-10
    assert abs(h["e_ccsd_corr"] - H2O_GOLD["ccsd_corr"]) < 5e-8      # reference CCSD threshold 1e-6
    _, b = fixture("butanol2_sto3g")
    assert abs(b["e_nuc"] - BUTANOL_GOLD["e_nuc"]) < 1e-11
    assert abs(b["e_scf"] - BUTANOL_GOLD["scf"]) < 5e-8
    assert abs(b["e_ccsd_corr"] - BUTANOL_GOLD["ccsd_corr"]) < 1e-6  # reference: threshold 1e-6, CD diagtol 1e-5


def test_spin_orbital_tensors_have_the_reference_symmetries():
    fx, _ = fixture("h2o_ccpvdz")
    T = pv.spin_orbital_tensors(fx)
    np.testing.assert_allclose(T["t2"], -T["t2"].transpose(1, 0, 2, 3), atol=1e-15)
    np.testing.assert_allclose(T["t2"], -T["t2"].transpose(0, 1, 3, 2), atol=1e-15)
    np.testing.assert_allclose(T["v2ijab"], -T["v2ijab"].transpose(1, 0, 2, 3), atol=1e-15)
    np.testing.assert_allclose(T["v2ijka"], -T["v2ijka"].transpose(1, 0, 2, 3), atol=1e-15)
    np.testing.assert_allclose(T["v2iabc"], -T["v2iabc"].transpose(0, 1, 3, 2), atol=1e-15)


def test_oracle_on_h2o_matches_reference_cpu_and_closed_form_and_is_tile_invariant():
    from oracle.oracle import Oracle, closed_form_energy
    fx, _ = fixture("h2o_ccpvdz")
    T = pv.spin_orbital_tensors(fx)
    orc = Oracle()
    e28 = orc.run(orc.tiles(5, 5, 19, 19, 28), T, True)
    assert abs(e28[0] - REFE["h2o_ccpvdz"]["E[T]"]) < 1e-13 and abs(e28[1] - REFE["h2o_ccpvdz"]["E(T)"]) < 1e-13
    e4 = orc.run(orc.tiles(5, 5, 19, 19, 4), T, True)
    assert abs(e4[0] - e28[0]) < 1e-13 and abs(e4[1] - e28[1]) < 1e-13
    c = closed_form_energy(5, 5, 19, 19, T, True)
    assert abs(c[0] - e28[0]) < 1e-12 and abs(c[1] - e28[1]) < 1e-12


def test_oracle_on_butanol2_reproduces_the_published_triples_corrections():
    """the reference's own CI golden for [T] and (T) -- through an integral code, SCF and CCSD that share
    nothing with the reference -- to 5e-9 Eh (observed 2e-9)"""
    from oracle.oracle import Oracle
    fx, _ = fixture("butanol2_sto3g")
    T = pv.spin_orbital_tensors(fx)
    orc = Oracle()
    sp = orc.tiles(21, 21, 14, 14, 40)
    assert orc.count_ops(sp, True) == BUTANOL_GOLD["total_num_ops"]
    e1, e2 = orc.run(sp, T, True)
    assert abs(e1 - BUTANOL_GOLD["[T]"]) < 5e-9 and abs(e2 - BUTANOL_GOLD["(T)"]) < 5e-9
    assert abs(e1 - REFE["butanol2_sto3g"]["E[T]"]) < 1e-13 and abs(e2 - REFE["butanol2_sto3g"]["E(T)"]) < 1e-13


def _gpu_energy(name, ts, **opts):
    from exachem_b200 import driver as drv
    fx, _ = fixture(name)
    T = pv.spin_orbital_tensors(fx)
    no, nv = int(fx["nocc"]), len(fx["eps"]) - int(fx["nocc"])
    sp = drv.setup_mo_space(no, no, nv, nv, ts)
    d = drv.CCSD_T_Fused_Driver(device=0, options=opts)
    e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, T["t1"], T["t2"],
                             {k: T[k] for k in ("v2ijab", "v2ijka", "v2iabc")}, T["evl"], 0.0, True)
    return e1, e2, d.last_stats


@pytest.mark.gpu
@pytest.mark.parametrize("ts", [28, 8, 5])
def test_gpu_h2o_matches_the_reference_cpu_energy(ts):
    e1, e2, st = _gpu_energy("h2o_ccpvdz", ts)
    r = REFE["h2o_ccpvdz"]
    assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL, (e1, e2, r)
    assert st["kernel_launches"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("ts", [40, 16])
def test_gpu_butanol2_matches_reference_cpu_and_published_golden(ts):
    e1, e2, st = _gpu_energy("butanol2_sto3g", ts)
    r = REFE["butanol2_sto3g"]
    assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL, (e1, e2, r)
    assert abs(e1 - BUTANOL_GOLD["[T]"]) < 5e-9 and abs(e2 - BUTANOL_GOLD["(T)"]) < 5e-9
    assert st["kernel_launches"] > 0


def _results_file(tmp_path, e1, e2, seconds):
    from exachem_b200 import results
    _, b = fixture("butanol2_sto3g")
    pt = results.ccsd_t_results(e1, e2, b["e_scf"], b["e_ccsd_corr"], seconds, seconds, BUTANOL_GOLD["total_num_ops"])
    path = os.path.join(tmp_path, "butanol2_pt.sto-3g.ccsd_t.json")
    results.write_json_data(path, {"SCF": {"conve": 1e-8}, "CC": {"threshold": 1e-6}}, b["e_scf"], b["e_ccsd_corr"], pt)
    return path, pt


def _check_like_the_ci_comparator(pt):
    """ci/scripts/compare_results.py:197-236: |ref - cur| <= CC.threshold (1e-6) on the six CCSD(T) energies"""
    gold_scf, gold_cc = BUTANOL_GOLD["scf"], BUTANOL_GOLD["ccsd_corr"]
    for key, corr in (("[T]Energies", BUTANOL_GOLD["[T]"]), ("(T)Energies", BUTANOL_GOLD["(T)"])):
        assert abs(pt[key]["correction"] - corr) <= 1e-6
        assert abs(pt[key]["correlation"] - (gold_cc + corr)) <= 1e-6
        assert abs(pt[key]["total"] - (gold_scf + gold_cc + corr)) <= 1e-6


def test_results_json_passes_the_reference_ci_comparator(tmp_path):
    r = REFE["butanol2_sto3g"]
    path, pt = _results_file(str(tmp_path), r["E[T]"], r["E(T)"], 30.9)
    assert set(pt) == {"[T]Energies", "(T)Energies", "performance"}
    assert set(pt["performance"]) == {"total_time", "gflops", "total_num_ops", "load_imbalance"}
    _check_like_the_ci_comparator(pt)
    script = "/root/reference/ci/scripts/compare_results.py"
    gold_dir = "/root/reference/ci/reference_output"
    if os.path.exists(script):          # in the build container: the reference's own comparator, unchanged, on our file
        import subprocess
        import sys
        # directory mode (its single-file mode trips over its own ref_notreq list): every other golden is reported
        # as "not available"; ours is compared for SCF, CCSD and the six CCSD(T) energies
        out = subprocess.run([sys.executable, script, gold_dir, os.path.dirname(path)], capture_output=True, text=True)
        assert "butanol2_pt.sto-3g.ccsd_t.json: Checking CCSD(T) results" in out.stdout, out.stdout + out.stderr
        assert "ERROR" not in out.stdout and "Traceback" not in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_gpu_results_json_passes_the_ci_thresholds(tmp_path):
    e1, e2, st = _gpu_energy("butanol2_sto3g", 40)
    _, pt = _results_file(str(tmp_path), e1, e2, st["seconds_total"])
    _check_like_the_ci_comparator(pt)
    assert pt["performance"]["total_num_ops"] == BUTANOL_GOLD["total_num_ops"] == st["counted_flops"]


LARGE_BENZENE = os.path.join(HERE, "golden", "_large", "benzene_ccpvdz.npz")


@pytest.mark.gpu
def test_gpu_benzene_real_amplitudes_match_the_reference_cpu_energy():
    """BASELINE.json configs[1] on REAL amplitudes: inputs/benzene.json (cc-pVDZ, 114 functions, O=21, V=93 per spin,
    ccsdt_tilesize 40) through tools/provider.  The ~200 MB fixture is git-ignored (tests/golden/make_benzene_large.py
    makes it; it travels with the snapshot); the reference's own CPU path on it takes hours on 8 cores, so its energies
    are stored in tests/golden/molecules_ref.json by tools/benzene_real.py --reference-cpu."""
    if not os.path.exists(LARGE_BENZENE):
        pytest.skip("tests/golden/_large/benzene_ccpvdz.npz not generated")
    from exachem_b200 import driver as drv
    e1, e2, st = _gpu_energy("_large/benzene_ccpvdz", 40)
    assert st["tasks_run"] == 28 and st["counted_flops"] == drv.count_ops(drv.setup_mo_space(21, 21, 93, 93, 40), True)
    assert -0.05 < e2 < -0.02 and e1 < e2           # benzene (T) is about -36 mEh; [T] overshoots it
    if "benzene_ccpvdz" in REFE:
        r = REFE["benzene_ccpvdz"]
        assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL, (e1, e2, r)


def test_unrestricted_task_list_gives_the_restricted_energy_on_h2o():
    """closed-shell data through the open-shell code path: is_restricted=false enumerates all four spin cases without the
    factor 2 (ccsd_t_fused_driver.hpp:383-395) and must reproduce the restricted energies"""
    from oracle.oracle import Oracle
    fx, _ = fixture("h2o_ccpvdz")
    T = pv.spin_orbital_tensors(fx)
    orc = Oracle()
    sp = orc.tiles(5, 5, 19, 19, 7)
    r = orc.run(sp, T, True)
    u = orc.run(sp, T, False)
    assert abs(r[0] - u[0]) < 1e-13 and abs(r[1] - u[1]) < 1e-13
    assert abs(r[1] - REFE["h2o_ccpvdz"]["E(T)"]) < 1e-13


@pytest.mark.gpu
def test_gpu_unrestricted_task_list_gives_the_restricted_energy_on_h2o():
    from exachem_b200 import driver as drv
    fx, _ = fixture("h2o_ccpvdz")
    T = pv.spin_orbital_tensors(fx)
    sp = drv.setup_mo_space(5, 5, 19, 19, 7)
    d = drv.CCSD_T_Fused_Driver(device=0)
    e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, T["t1"], T["t2"], {k: T[k] for k in ("v2ijab", "v2ijka", "v2iabc")},
                             T["evl"], 0.0, False)
    r = REFE["h2o_ccpvdz"]
    assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL


def _ch2():
    fx = np.load(os.path.join(HERE, "golden", "ch2_triplet_321g.npz"))
    T = {k: fx[k] for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")}
    return T, json.loads(str(fx["summary"]))


def test_open_shell_ch2_triplet_oracle_reference_and_unequal_tile_counts():
    """row f5: triplet CH2 (UHF, 5 alpha / 3 beta occupied) -- one tile per spin block reproduces the reference CPU path;
    tilings with DIFFERENT numbers of alpha and beta tiles (which the reference's half/half k_spin rule, ccsd_t.cpp:245-249,
    mislabels) give the same energy and the closed form of the 27 equations"""
    from oracle.oracle import Oracle, closed_form_energy
    T, info = _ch2()
    na, nb, n = info["n_occ_alpha"], info["n_occ_beta"], info["nbf"]
    assert abs(info["s_squared"] - 2.0) < 0.05                       # a triplet
    orc = Oracle()
    r = REFE["ch2_triplet_321g"]
    e40 = orc.run(orc.tiles(na, nb, n - na, n - nb, 40), T, False)
    assert abs(e40[0] - r["E[T]"]) < 1e-14 and abs(e40[1] - r["E(T)"]) < 1e-14
    c = closed_form_energy(na, nb, n - na, n - nb, T, False)
    for ts in (4, 3, 2):                                             # occ alpha/beta tiles 2/1, 2/1, 3/2
        sp = orc.tiles(na, nb, n - na, n - nb, ts)
        assert sp.noa != sp.nob
        assert list(sp.k_spin[:sp.noab]) == [1] * sp.noa + [2] * sp.nob
        e = orc.run(sp, T, False)
        assert abs(e[0] - c[0]) < 1e-14 and abs(e[1] - c[1]) < 1e-14 and abs(e[1] - r["E(T)"]) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("ts", [40, 4, 3, 2])
def test_gpu_open_shell_ch2_triplet(ts):
    from exachem_b200 import driver as drv
    T, info = _ch2()
    na, nb, n = info["n_occ_alpha"], info["n_occ_beta"], info["nbf"]
    sp = drv.setup_mo_space(na, nb, n - na, n - nb, ts)
    d = drv.CCSD_T_Fused_Driver(device=0)
    e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, T["t1"], T["t2"], {k: T[k] for k in ("v2ijab", "v2ijka", "v2iabc")},
                             T["evl"], 0.0, False)
    r = REFE["ch2_triplet_321g"]
    assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL, (ts, e1, e2, r)


CH4_GOLD = {"[T]": -0.00802036907441595, "(T)": -0.007817304224704037, "total_num_ops": 37432196256}
LARGE_CH4 = os.path.join(HERE, "golden", "_large", "ch4_def2tzvp_uhf.npz")


@pytest.mark.gpu
@pytest.mark.parametrize("exec_tilesize", [0, -1])
def test_gpu_ch4_triplet_uhf_reproduces_the_published_open_shell_golden(exec_tilesize):
    """The reference's own OPEN-SHELL CI golden (ci/reference_output/ch4.def2-tzvp.ccsd_t.json; inputs/ci/ch4.json: UHF triplet,
    def2-TZVP, 6 alpha / 4 beta electrons, 49 / 51 virtuals, ccsdt_tilesize 28 -> unequal alpha / beta tile extents) on the GPU,
    with is_restricted = false: [T] and (T) to 2e-8 Eh of the published values (the reference stopped its CCSD at 1e-6;
    observed 7e-9) and to 1e-9 Eh of the oracle on the same amplitudes; total_num_ops exact.  On the caller's tiles and on
    the automatic execution tiling (28 -> 40: the ragged 21 / 23 remainders disappear).  The ~90 MB fixture is git-ignored
    (tests/golden/make_ch4_large.py) and travels with the snapshot."""
    if not os.path.exists(LARGE_CH4) or "ch4_def2tzvp_uhf" not in REFE:
        pytest.skip("tests/golden/_large/ch4_def2tzvp_uhf.npz not generated")
    from exachem_b200 import driver as drv
    fx = np.load(LARGE_CH4)
    T = {k: fx[k] for k in ("t1", "t2", "v2ijab", "v2ijka", "v2iabc", "evl")}
    r = REFE["ch4_def2tzvp_uhf"]
    na, nb, n = r["n_occ_alpha"], r["n_occ_beta"], r["nbf"]
    sp = drv.setup_mo_space(na, nb, n - na, n - nb, 28)
    assert drv.count_ops(sp, False) == CH4_GOLD["total_num_ops"]
    d = drv.CCSD_T_Fused_Driver(device=0, options={"exec_tilesize": exec_tilesize})
    e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, T["t1"], T["t2"], {k: T[k] for k in ("v2ijab", "v2ijka", "v2iabc")},
                             T["evl"], 0.0, False)
    assert abs(e1 - r["E[T]"]) <= ATOL and abs(e2 - r["E(T)"]) <= ATOL, (e1, e2, r)
    assert abs(e1 - CH4_GOLD["[T]"]) < 2e-8 and abs(e2 - CH4_GOLD["(T)"]) < 2e-8
    assert d.last_stats["kernel_launches"] > 0

