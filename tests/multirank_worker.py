"""One rank of the multi-process tests of the C++ drop-in header (tests/test_multirank_cpp.py): the ranks are separate
processes that meet in the TAMM stand-in's shared-memory ProcGroup (TAMM_SHIM_RANK / _SIZE / _KEY in the environment).
  argv: mode fixture-name result.json [harness-flavour]
  mode selftest: host-only (barrier, broadcast, shared task counter);  mode execute: CCSD_T_Fused_Driver::execute on a GPU."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exachem_b200 import _lib, driver as drv, synthetic as syn  # noqa: E402

mode, name, out_path = sys.argv[1:4]
flavour = sys.argv[4] if len(sys.argv) > 4 else ""
so = os.path.join(ROOT, "tests", "cpp", "_build", f"libadapter_test{flavour}.so")
_lib.load()
L = C.CDLL(so)
L.adapter_last_error.restype = C.c_char_p
rank = int(os.environ["TAMM_SHIM_RANK"])
if mode == "selftest":
    out = np.zeros(5, np.int64)
    L.adapter_multirank_selftest.argtypes = [C.c_int64, C.c_char_p, _lib._i64p]
    rc = L.adapter_multirank_selftest(int(name), f"/ccsdt_test_{os.environ['TAMM_SHIM_KEY']}".encode(), out.ctypes.data_as(_lib._i64p))
    json.dump({"rc": rc, "out": out.tolist(), "err": L.adapter_last_error().decode()}, open(out_path, "w"))
    sys.exit(0)

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_small.json")))
g = GOLD[name]
if os.environ.get("WORKER_DEVICE"):
    import torch
    torch.cuda.set_device(int(os.environ["WORKER_DEVICE"]))     # the rank's current CUDA device, as TAMM would bind it
sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
kr, ks = np.ascontiguousarray(sp.k_range, np.int64), np.ascontiguousarray(sp.k_spin, np.int32)
out, gets, st = np.zeros(4), np.zeros(5, np.int64), _lib.Stats()
a = {k: np.ascontiguousarray(T[k]) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")}
L.adapter_ccsdt_execute.restype = C.c_int
L.adapter_ccsdt_execute.argtypes = [C.c_int] * 4 + [_lib._i64p, _lib._i32p] + [_lib._dp] * 6 + [C.c_int, C.c_int, _lib._dp, _lib._i64p,
                                                                                               C.POINTER(C.c_longdouble), C.POINTER(_lib.Stats)]
rc = L.adapter_ccsdt_execute(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p),
                             *[a[k].ctypes.data_as(_lib._dp) for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")],
                             int(g["restricted"]), g["tilesize"], out.ctypes.data_as(_lib._dp), gets.ctypes.data_as(_lib._i64p), None,
                             C.byref(st))
json.dump({"rc": rc, "err": L.adapter_last_error().decode(), "e1": out[0], "e2": out[1], "tasks_run": st.tasks_run,
           "blocks_fetched": st.blocks_fetched, "blocks_from_peers": st.blocks_from_peers, "blocks_evicted": st.blocks_evicted, "gets": int(gets.sum()), "rank": rank}, open(out_path, "w"))
