"""Times ccsdt_put_cholesky (V2 tensors from Cholesky vectors on the device) at the benzene cc-pVDZ size.
usage: python tools/time_put_cholesky.py [O V NCV]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from exachem_b200 import driver as drv, synthetic as syn

o, v, ncv = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (21, 93, 900)
sp = drv.setup_mo_space(o, o, v, v, 40)
n = 2 * (o + v)
rng = np.random.default_rng(0)
L = rng.uniform(-1, 1, (n, n, ncv))
ctx = drv.Context(0)
ctx.set_space(sp, syn.Orbitals(o, o, v, v).orbital_energies(), True)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.put_cholesky(L)
    dt = time.perf_counter() - t0
    O, V = 2 * o, 2 * v
    flops = 2.0 * ncv * ((O * V) ** 2 + O * O * O * V + O * V * V * V)
    print(f"put_cholesky O={O} V={V} ncv={ncv}: {dt * 1e3:.1f} ms  (H2D {L.nbytes / 1e9:.2f} GB, {flops / 1e12:.2f} TFLOP of DGEMM, "
          f"{(O * O * V * V + O ** 3 * V + O * V ** 3) * 8 / 1e9:.2f} GB of V2 written)")
ctx.close()
