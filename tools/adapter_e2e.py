"""End-to-end through the C++ drop-in header (include/ccsd_t_fused_driver_b200.hpp, compiled against the TAMM
stand-in by tests/cpp): CCSD_T_Fused_Driver<double>::execute on HOST tensors of the benzene shape, every block
pulled through Tensor::get -> fetch callback -> H2D.  Prints the wall time of execute, the counted TFLOP/s and
the difference to the device-generated run.   usage: python tools/adapter_e2e.py [noa nva ts]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import _lib, driver as drv, synthetic as syn  # noqa: E402
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("test_cpp_adapter", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "test_cpp_adapter.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
harness = _mod.harness

no, nv, ts = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (21, 93, 40)
L = _lib.load()
H = harness(vector_get=bool(int(os.environ.get('ADAPTER_VECTOR_GET', '0'))))
sp = drv.setup_mo_space(no, no, nv, nv, ts)
O, V = 2 * no, 2 * nv
dims = {drv.T1: (V, O, 1, 1), drv.T2: (V, V, O, O), drv.V_IJAB: (O, O, V, V), drv.V_IJKA: (O, O, O, V), drv.V_IABC: (O, V, V, V)}
host = {}
for tid, d in dims.items():
    buf = np.zeros(int(np.prod(d)))
    lo, nn = np.zeros(4, np.int64), np.array(d, np.int64)
    assert L.ccsdt_synth_block(0, 1234, tid, no, no, nv, nv, lo.ctypes.data_as(_lib._i64p), nn.ctypes.data_as(_lib._i64p),
                               buf.ctypes.data_as(_lib._dp)) == 0
    host[tid] = buf
evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
kr, ks = np.ascontiguousarray(sp.k_range, np.int64), np.ascontiguousarray(sp.k_spin, np.int32)
out, gets, ops, st = np.zeros(4), np.zeros(5, np.int64), C.c_longdouble(0), _lib.Stats()
for rep in range(int(os.environ.get('ADAPTER_REPS', '4'))):
    t0 = time.perf_counter()
    rc = H.adapter_ccsdt_execute(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p),
                                 evl.ctypes.data_as(_lib._dp), *[host[t].ctypes.data_as(_lib._dp) for t in range(5)], 1, ts,
                                 out.ctypes.data_as(_lib._dp), gets.ctypes.data_as(_lib._i64p), C.byref(ops), C.byref(st))
    dt = time.perf_counter() - t0
    assert rc == 0, H.adapter_last_error()
    print(f"adapter execute #{rep}: {dt:.3f} s wall, {float(ops.value) / dt / 1e12:.2f} TFLOP/s counted end to end; "
          f"ccsdt_run {st.seconds_total:.3f} s, kernel {st.seconds_kernel:.3f} s, staging {st.seconds_staging:.3f} s, {st.blocks_fetched} blocks / "
          f"{st.h2d_bytes / 1e9:.2f} GB fetched through Tensor::get ({st.seconds_fetch:.3f} s inside the callback, host waited {st.seconds_host_wait:.3f} s "
          f"for the GPU), E[T] {out[0]:.12e} E(T) {out[1]:.12e}")
ctx = drv.Context(0)
ctx.set_space(sp, evl, True)
ctx.set_synthetic(1234)
e1, e2, st2, _ = ctx.run()
ctx.close()
print(f"device-generated run: E[T] {e1:.12e} E(T) {e2:.12e}; |diff| {abs(e1 - out[0]):.3e} {abs(e2 - out[1]):.3e}")
