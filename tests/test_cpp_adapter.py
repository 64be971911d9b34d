"""The C++ drop-in header include/ccsd_t_fused_driver_b200.hpp (class CCSD_T_Fused_Driver<T> with the
reference's execute / calculate_performance_ops signatures, ccsd_t_fused_driver.hpp:73-87), compiled
against the TAMM stand-in by tests/cpp/Makefile and called exactly as ccsd_t.cpp:253-256,298-301 does.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from exachem_b200 import _lib, driver as drv, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "_build", "libadapter_test.so")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_small.json")))
_dp, _i64p, _i32p = _lib._dp, _lib._i64p, _lib._i32p


def harness(vector_get=False):
    """vector_get: the harness built with the std::vector flavour of Tensor::get instead of the span flavour"""
    so = SO.replace("libadapter_test.so", "libadapter_test_vec.so") if vector_get else SO
    if not os.path.exists(so):   # on the GPU box the prebuilt file travels with the snapshot
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    _lib.load()
    L = C.CDLL(so)
    L.adapter_last_error.restype = C.c_char_p
    L.adapter_ccsdt_execute.restype = C.c_int
    L.adapter_ccsdt_execute.argtypes = [C.c_int] * 4 + [_i64p, _i32p] + [_dp] * 6 + [C.c_int, C.c_int, _dp, _i64p,
                                                                                    C.POINTER(C.c_longdouble),
                                                                                    C.POINTER(_lib.Stats)]
    L.adapter_ccsdt_count_ops.restype = C.c_int
    L.adapter_ccsdt_count_ops.argtypes = [C.c_int] * 4 + [_i64p, _i32p, C.c_int, C.POINTER(C.c_longdouble)]
    return L


def _space_args(sp):
    kr = np.ascontiguousarray(sp.k_range, np.int64)
    ks = np.ascontiguousarray(sp.k_spin, np.int32)
    return kr, ks


def test_adapter_compiles_and_counts_ops_like_the_reference_goldens():
    """host-only: calculate_performance_ops through the C++ adapter against the CI goldens' total_num_ops"""
    L = harness()
    for oa, ob, va, vb, ts, r, gold in [(21, 21, 14, 14, 40, True, 30952040112), (6, 4, 49, 51, 28, False, 37432196256),
                                        (9, 9, 36, 36, 40, True, 52991044992),
                                        (29, 29, 103, 103, 40, True, 62789556886348)]:
        sp = drv.setup_mo_space(oa, ob, va, vb, ts)
        kr, ks = _space_args(sp)
        v = C.c_longdouble(0)
        assert L.adapter_ccsdt_count_ops(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_i64p),
                                         ks.ctypes.data_as(_i32p), int(r), C.byref(v)) == 0
        assert int(v.value) == gold


def test_adapter_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = harness()
    g = GOLD[sorted(GOLD)[0]]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    kr, ks = _space_args(sp)
    out = np.zeros(4)
    a = {k: np.ascontiguousarray(T[k]) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")}
    rc = L.adapter_ccsdt_execute(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_i64p), ks.ctypes.data_as(_i32p),
                                 *[a[k].ctypes.data_as(_dp) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")],
                                 int(g["restricted"]), g["tilesize"], out.ctypes.data_as(_dp), None, None, None)
    assert rc == 1 and b"no CPU fallback" in L.adapter_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("vector_get", [False, True])
@pytest.mark.parametrize("exec_tilesize", [0, -1])
@pytest.mark.parametrize("name", sorted(GOLD))
def test_adapter_execute_matches_reference_fixture(name, vector_get, exec_tilesize, monkeypatch):
    """same call as ccsd_t.cpp:253-256 on the tensors the committed reference fixture was made from; both flavours of
    Tensor::get (span: blocks land in the library's pinned buffer; std::vector: one extra host copy); on the caller's tiles
    (CCSDT_B200_EXEC_TILESIZE=0) and on the adapter's default, the automatic execution tiling"""
    L = harness(vector_get)
    g = GOLD[name]
    monkeypatch.setenv("CCSDT_B200_EXEC_TILESIZE", str(exec_tilesize))
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    kr, ks = _space_args(sp)
    out, gets, ops, st = np.zeros(4), np.zeros(5, np.int64), C.c_longdouble(0), _lib.Stats()
    a = {k: np.ascontiguousarray(T[k]) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")}
    rc = L.adapter_ccsdt_execute(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_i64p), ks.ctypes.data_as(_i32p),
                                 *[a[k].ctypes.data_as(_dp) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")],
                                 int(g["restricted"]), g["tilesize"], out.ctypes.data_as(_dp),
                                 gets.ctypes.data_as(_i64p), C.byref(ops), C.byref(st))
    assert rc == 0, L.adapter_last_error()
    assert abs(out[0] - float(g["energy1"])) <= 1e-9 and abs(out[1] - float(g["energy2"])) <= 1e-9   # 1e-9 Eh
    assert int(ops.value) == g["total_num_ops"]
    assert st.kernel_launches > 0
    if exec_tilesize == 0:                              # the caller's tiles: the reference's task list, task for task
        assert st.tasks_run == len(g["tasks"])
    assert gets.sum() == st.blocks_fetched > 0          # every block fetched through Tensor::get exactly once
