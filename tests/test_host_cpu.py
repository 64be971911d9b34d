"""CPU-side tests of the product library: it loads, exports every symbol include/ccsdt_b200.h
declares, its host logic (tiling, enumeration, enabled terms, flop count, rank split) is bit-exact
against the oracle, and it refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from exachem_b200 import _lib, driver as drv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = [  # n_occ_a, n_occ_b, n_vir_a, n_vir_b, tilesize, restricted
    (4, 4, 6, 6, 3, True), (4, 4, 6, 6, 2, True), (3, 3, 9, 9, 4, True), (5, 5, 11, 11, 8, True),
    (3, 2, 5, 6, 3, False), (4, 4, 6, 6, 4, False), (5, 5, 19, 19, 28, True), (21, 21, 93, 93, 40, True),
    (6, 4, 49, 51, 28, False), (29, 29, 103, 103, 40, True), (1, 1, 1, 1, 1, True), (7, 7, 3, 3, 2, False),
]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ccsdt_b200.h")).read()
    declared = set(re.findall(r"CCSDT_API\s+[\w\s\*]+?\b(ccsdt_\w+)\(", hdr))
    assert len(declared) >= 20
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_host_logic_matches_oracle(orc, cfg):
    oa, ob, va, vb, ts, restricted = cfg
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    assert np.array_equal(sp.k_range, osp.k_range) and np.array_equal(sp.k_spin, osp.k_spin)
    assert (sp.noa, sp.nob, sp.nva, sp.nvb) == (osp.noa, osp.nob, osp.nva, osp.nvb)
    tasks, fac, n_outer = drv.enumerate_tasks(sp, restricted)
    otasks, ofac, on_outer = orc.enumerate(osp, restricted)
    assert np.array_equal(tasks, otasks)          # bit-exact, in the reference's order
    assert np.array_equal(fac, ofac) and n_outer == on_outer
    assert drv.count_ops(sp, restricted) == orc.count_ops(osp, restricted)
    for t in tasks[:: max(1, len(tasks) // 200)]:
        s1, d1, d2 = drv.task_terms(sp, restricted, t)
        os1, od1, od2, _ = orc.task_exec(osp, restricted, t)
        assert np.array_equal(s1, os1 >= 0) and np.array_equal(d1, od1 >= 0) and np.array_equal(d2, od2 >= 0)


def test_flop_count_known_answers():
    """the reference CI goldens' total_num_ops (see tests/test_oracle.py for sources)"""
    for oa, ob, va, vb, ts, r, gold in [(21, 21, 14, 14, 40, True, 30952040112), (6, 4, 49, 51, 28, False, 37432196256),
                                        (9, 9, 36, 36, 40, True, 52991044992),
                                        (29, 29, 103, 103, 40, True, 62789556886348),
                                        (146, 146, 591, 591, 40, True, 3088106365505672192)]:
        assert drv.count_ops(drv.setup_mo_space(oa, ob, va, vb, ts), r) == gold


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_is_balanced_and_complete(nranks):
    sp = drv.setup_mo_space(20, 20, 60, 60, 12)
    own = drv.partition(sp, True, nranks)
    tasks, _, _ = drv.enumerate_tasks(sp, True)
    assert len(own) == len(tasks) and own.min() >= 0 and own.max() < nranks
    ext = sp.k_range
    cost = np.array([np.prod(ext[t[:6]].astype(float)) for t in tasks])
    loads = np.array([cost[own == r].sum() for r in range(nranks)])
    assert loads.min() > 0.8 * loads.max()


def test_check_memory_bound():
    L = _lib.load()
    import ctypes as C
    need = C.c_size_t(0)
    assert L.ccsdt_check_memory(40, 500, 180 * 10**9, C.byref(need)) == 0
    assert need.value == int(9 * (40**2 + 40**4 + 4 * 500 * 40**3) * 8)
    assert L.ccsdt_check_memory(40, 500, 10**9, C.byref(need)) == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(drv.CcsdtError, match="no CPU fallback"):
        drv.Context(0)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under exachem_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "exachem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt \
                    and "ccsdt_oracle" not in txt and "libccsdt_ref" not in txt, os.path.join(dirpath, f)


@pytest.mark.parametrize("sym", range(16))
def test_box_weights_count_every_box_once(sym):
    """symmetry reduction of the fused kernel (options.symmetry): over the boxes of a tile whose coinciding
    indices have equal box counts, the weights of the surviving boxes add up to the number of boxes"""
    import ctypes as C
    import itertools
    L = _lib.load()
    nh, nh3, np_, np6 = 4, 3, 3, 2
    # coinciding indices must have the same number of boxes
    n = [nh, nh if sym & 1 else 5, 0, np_, np_ if sym & 4 else 4, 0]
    n[2] = n[1] if sym & 2 else nh3
    n[5] = n[4] if sym & 8 else np6
    total, kept = 0, 0
    for box in itertools.product(*[range(k) for k in n]):
        b = (C.c_int32 * 6)(*box)
        w = L.ccsdt_box_weight(sym, b)
        total += w
        kept += w > 0
        assert w in (0, 1, 2, 3, 6, 4, 9, 12, 18, 36)
    assert total == int(np.prod(n))
    if sym:
        assert kept < int(np.prod(n))


def test_box_weight_skips_the_triple_hole_diagonal():
    import ctypes as C
    import itertools
    L = _lib.load()
    n = [5, 5, 5, 2, 3, 2]
    total = sum(L.ccsdt_box_weight(3 | 16, (C.c_int32 * 6)(*b)) for b in itertools.product(*[range(k) for k in n]))
    assert total == int(np.prod(n)) - 5 * 2 * 3 * 2       # the 5 diagonal hole boxes of every particle box are dropped


def test_memory_estimate_matches_the_survey_sizes():
    """ccsdt_estimate_memory (the counterpart of the memory summary of ccsd_t.cpp:95-152): tensor sizes of SURVEY.md 8d at
    (60,500) -- T2 and v2ijab 43 GB, v2ijka 5.2 GB, v2iabc 360 GB, 2.7 GB of blocks per task -- and the panel pools"""
    import ctypes as C
    L = _lib.load()

    def est(o, v, ts, target):
        sp = drv.setup_mo_space(o, o, v, v, ts)
        m = _lib.MemoryEstimate()
        kr, ks = np.ascontiguousarray(sp.k_range, np.int64), np.ascontiguousarray(sp.k_spin, np.int32)
        assert L.ccsdt_estimate_memory(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p),
                                       ks.ctypes.data_as(_lib._i32p), target, C.byref(m)) == 0
        return m

    m = est(60, 500, 32, 0)
    assert [round(x / 1e9, 1) for x in m.tensor_bytes] == [0.0, 43.2, 43.2, 5.2, 360.0]
    assert m.tensor_bytes[4] == 6 * 60 * 500 ** 3 * 8
    assert abs(m.task_block_bytes / 1e9 - 2.7) < 0.1
    # two staging buffers x 9 slots x (HPP 32*32*32 + HHP 32*32*32) x (Kp 64 + 512) doubles
    assert m.panel_bytes == 2 * 8 * 9 * (32 ** 3 + 32 ** 3) * (64 + 512)
    assert m.minimum_bytes == m.panel_bytes + m.s1_bytes + 2 * m.task_block_bytes
    caf_own, caf_exec = est(51, 195, 28, 0), est(51, 195, 28, -1)
    assert (caf_own.exec_max_particle_tile, caf_exec.exec_max_particle_tile) == (28, 40)
    assert caf_own.tensor_bytes[:] == caf_exec.tensor_bytes[:]           # the tensors do not depend on the execution tiling

