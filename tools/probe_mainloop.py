"""DMMA issue/mainloop microbenchmark: TFLOP/s vs (tiles per warp, warps per CTA, CTAs per SM)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import _lib
L = _lib.load()
tf = C.c_double(0); ms = C.c_double(0)
for name, flag in (("dfma", 0), ("dmma", 1)):
    L.ccsdt_probe_fp64_peak(0, flag, 20000, C.byref(tf), C.byref(ms)); print(f"peak {name}: {tf.value:.2f} TF/s")
for (ta, tb) in ((4, 4), (2, 4), (2, 2), (4, 8)):
    for wpc, cps in ((4, 1), (8, 1), (12, 1), (16, 1), (4, 2), (8, 2), (4, 3), (4, 4), (8, 3), (8, 4), (16, 2), (16,3), (16, 4)):
        rc = L.ccsdt_probe_mainloop(0, ta, tb, wpc, cps, 4000, C.byref(tf))
        print(f"tiles {ta}x{tb} warps/cta {wpc:2d} ctas/sm {cps} -> warps/SM {wpc*cps:2d}: " + (f"{tf.value:6.2f} TF/s" if rc == 0 else "n/a"))
