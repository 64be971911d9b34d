#!/usr/bin/env python
"""(T) of benzene cc-pVDZ on REAL converged amplitudes (tests/golden/_large/benzene_ccpvdz.npz, see
tests/golden/make_benzene_large.py): this repo's GPU path next to the reference's own implementations.

  python tools/benzene_real.py --gpu            # ours (symmetry on and off) + the reference's GPU kernel K1 (oracle/_ref), on a B200
  python tools/benzene_real.py --reference-cpu  # the reference's CPU path (oracle/_ref, hours on 8 cores); stores the energies
                                                # in tests/golden/molecules_ref.json (build container only)
TEST/MEASUREMENT TOOL: the only product code it calls is the public driver.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FIX = os.path.join(ROOT, "tests", "golden", "_large", "benzene_ccpvdz.npz")
REFJ = os.path.join(ROOT, "tests", "golden", "molecules_ref.json")
TS = 40


def tensors():
    from tools.provider import provider as pv
    fx = np.load(FIX)
    no, nv = int(fx["nocc"]), len(fx["eps"]) - int(fx["nocc"])
    return pv.spin_orbital_tensors(fx), no, nv, json.loads(str(fx["summary"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--reference-cpu", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "benzene_real.json"))
    args = ap.parse_args()
    T, no, nv, info = tensors()
    res = {"molecule": "benzene cc-pVDZ (inputs/benzene.json)", "provider": info}
    if args.reference_cpu:
        from oracle.oracle import Oracle, Reference
        orc, ref = Oracle(), Reference()
        sp = orc.tiles(no, no, nv, nv, TS)
        t0 = time.time()
        e, trace = ref.execute(sp, T, True, tilesize=TS)
        dt = time.time() - t0
        allr = json.load(open(REFJ))
        allr["benzene_ccpvdz"] = {"ccsdt_tilesize": TS, "E[T]": float(e[0]), "E(T)": float(e[1]), "kernel_tasks": int(len(trace)),
                                  "seconds": dt, "threads": ref.num_threads(),
                                  "source": "CCSD_T_Fused_Driver<double>::execute + total_fused_ccsd_t_cpu (oracle/_ref) on "
                                            "tests/golden/_large/benzene_ccpvdz.npz"}
        json.dump(allr, open(REFJ, "w"), indent=1)
        print(allr["benzene_ccpvdz"])
        return
    if args.gpu:
        from exachem_b200 import driver as drv
        from oracle.oracle import REF_GPU_SO, Oracle, ReferenceGPU
        sp = drv.setup_mo_space(no, no, nv, nv, TS)
        for sym in (1, 0):
            d = drv.CCSD_T_Fused_Driver(device=0, options={"symmetry": sym})
            for _ in range(2):
                e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, T["t1"], T["t2"],
                                         {k: T[k] for k in ("v2ijab", "v2ijka", "v2iabc")}, T["evl"], 0.0, True)
            st = d.last_stats
            res[f"ours_symmetry_{sym}"] = {"E[T]": e1, "E(T)": e2, "kernel_ms": st["seconds_kernel"] * 1e3,
                                           "total_ms": st["seconds_total"] * 1e3, "counted_flops": st["counted_flops"],
                                           "evaluated_flops": st["evaluated_flops"],
                                           "tflops_counted_kernel": st["counted_flops"] / st["seconds_kernel"] / 1e12}
            print(f"ours symmetry={sym}: E[T] {e1:.15f} E(T) {e2:.15f} kernel {st['seconds_kernel'] * 1e3:.1f} ms", flush=True)
        if os.path.exists(REF_GPU_SO["tc"]):
            orc = Oracle()
            ref = ReferenceGPU("tc")
            out, trace = ref.execute(orc.tiles(no, no, nv, nv, TS), T, True, tilesize=TS)
            res["reference_gpu_K1"] = {"E[T]": float(out[0]), "E(T)": float(out[1]), "kernel_ms": float(trace[:, 7].sum()),
                                       "kernel": "fully_fused_kernel_ccsd_t_nvidia_tc_fp64, unmodified, sm_100a"}
            print(f"reference K1: E[T] {out[0]:.15f} E(T) {out[1]:.15f} kernel {trace[:, 7].sum():.1f} ms", flush=True)
            res["max_abs_diff_ours_vs_reference_gpu"] = max(abs(res["ours_symmetry_1"]["E[T]"] - out[0]),
                                                            abs(res["ours_symmetry_1"]["E(T)"] - out[1]))
        if os.path.exists(REFJ) and "benzene_ccpvdz" in json.load(open(REFJ)):
            r = json.load(open(REFJ))["benzene_ccpvdz"]
            res["reference_cpu"] = r
            res["max_abs_diff_ours_vs_reference_cpu"] = max(abs(res["ours_symmetry_1"]["E[T]"] - r["E[T]"]),
                                                            abs(res["ours_symmetry_1"]["E(T)"] - r["E(T)"]))
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
        print({k: v for k, v in res.items() if k.startswith("max_abs")})


if __name__ == "__main__":
    main()
