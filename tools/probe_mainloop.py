"""DMMA issue/mainloop microbenchmark: TFLOP/s of (a) register-resident DFMA / DMMA loops, (b) an unrolled
load-then-compute loop (mode 4x4), (c) the product kernel's rolled k-step loop without barriers (mode 1) and
with the per-slab mbarrier traffic (mode 2), at several warps-per-SM settings."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exachem_b200 import _lib
L = _lib.load()
tf = C.c_double(0); ms = C.c_double(0)
for name, flag in (("dfma", 0), ("dmma", 1)):
    L.ccsdt_probe_fp64_peak(0, flag, 20000, C.byref(tf), C.byref(ms)); print(f"peak {name}: {tf.value:.2f} TF/s")
for (ta, tb, label) in ((4, 4, "unrolled 4x4"), (1, 0, "rolled loop"), (2, 0, "rolled loop + mbarriers"), (3, 0, "same loop unrolled x2"), (5, 0, "same loop unrolled x4")):
    for wpc, cps in ((4, 1), (4, 2), (4, 3), (4, 4)):
        rc = L.ccsdt_probe_mainloop(0, ta, tb, wpc, cps, 4000, C.byref(tf))
        print(f"{label:26s} warps/cta {wpc:2d} ctas/sm {cps} -> warps/SM {wpc*cps:2d}: " + (f"{tf.value:6.2f} TF/s" if rc == 0 else f"n/a ({L.ccsdt_last_error(None).decode()})"))
