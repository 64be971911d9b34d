"""The bench's e2e step alone (1 GPU): CCSD_T_Fused_Driver<double>::execute of the C++ header on the host block table of the
(60,500) ts32 sample.  CCSDT_B200_VERBOSE=2 prints the per-task time line of every call.  usage: python tools/table_e2e.py [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from exachem_b200 import _lib, driver as drv  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = bench.SYNTH
H = bench.Harness()
sp = drv.setup_mo_space(w["noa"], w["nob"], w["nva"], w["nvb"], w["ts"])
evl = bench.orbital_energies(w)
tasks, _, _ = drv.enumerate_tasks(sp, True)
ids, sel = bench.select_tasks(tasks, ("triples", bench.N_TRIPLES), 1)
kr, ks = np.ascontiguousarray(sp.k_range, np.int64), np.ascontiguousarray(sp.k_spin, np.int32)
table = H.L.adapter_table_open(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p),
                               evl.ctypes.data_as(_lib._dp), bench.SEED, 0)
cnt = np.array([sp.noa, sp.nob, sp.nva, sp.nvb], np.int32)
opt = H.options()
out4, gets, st = np.zeros(4), np.zeros(5, np.int64), _lib.Stats()
sub = np.ascontiguousarray(ids, np.int64)
for rep in range(reps + 1):
    t0 = time.perf_counter()
    rc = H.L.adapter_table_execute(table, cnt.ctypes.data_as(_lib._i32p), 1, w["ts"], sub.ctypes.data_as(_lib._i64p), len(sub), C.byref(opt),
                                   out4.ctypes.data_as(_lib._dp), C.byref(st), gets.ctypes.data_as(_lib._i64p))
    dt = time.perf_counter() - t0
    assert rc == 0, H.L.adapter_last_error()
    print(f"execute #{rep}{' (fills the host table)' if rep == 0 else ''}: {dt:.3f} s wall; run {st.seconds_total:.3f} s, kernel {st.seconds_kernel:.3f} s, "
          f"Tensor::get {st.seconds_fetch:.3f} s, host waited {st.seconds_host_wait:.3f} s, {st.blocks_fetched} blocks {st.h2d_bytes / 1e9:.1f} GB, "
          f"evicted {st.blocks_evicted}", flush=True)
H.L.adapter_table_close(table)
