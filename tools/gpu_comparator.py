#!/usr/bin/env python
"""On-box GPU comparator: the reference's OWN GPU kernels -- K1
`fully_fused_kernel_ccsd_t_nvidia_tc_fp64` (sm_80 DMMA, ccsd_t_all_fused_gpu.cu:132) and K2
`revised_jk_ccsd_t_fully_fused_kernel` (FMA, ccsd_t_all_fused_nontcCuda_Hip_Sycl.cpp:95) -- compiled
unmodified for sm_100a (oracle/_ref/libccsdt_refgpu_{tc,fma}.so) and driven by the reference's own GPU
task function, next to this repo's fused_t_dmma_kernel, on the same B200, the same procedural tensors
and the same kernel tasks.  TEST/MEASUREMENT TOOL: the only place oracle/_ref GPU code runs.

  python tools/gpu_comparator.py [--out gpurun_out/comparator.json] [--skip-fma]

Kernel times: reference = CUDA events around the reference launcher on the reference's stream (its five
constant-memory uploads + the kernel); ours = Stats.seconds_kernel (CUDA events around the fused kernel).
The reference's host staging runs against the naive HPTT stand-in, so only kernel times are compared.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEED = 1234


def orbital_energies(noa, nob, nva, nvb):
    from exachem_b200 import synthetic as syn
    return syn.Orbitals(noa, nob, nva, nvb).orbital_energies()


def task_flops(orc, osp, task):
    s1, d1, d2, _ = orc.task_exec(osp, True, task)
    base = 2.0 * float(np.prod(osp.k_range[task[:6]].astype(float)))
    f = base * int((s1 >= 0).sum())
    for h7 in range(osp.noab):
        f += base * int((d1[9 * h7:9 * h7 + 9] >= 0).sum()) * int(osp.k_range[h7])
    for p7 in range(osp.nvab):
        f += base * int((d2[9 * p7:9 * p7 + 9] >= 0).sum()) * int(osp.k_range[osp.noab + p7])
    return f


def run_shape(name, noa, nva, ts, ntasks, kinds):
    """first `ntasks` kernel tasks (all when ntasks < 0) of a closed-shell (noa, nva) problem at tile ts"""
    from exachem_b200 import driver as drv
    from oracle.oracle import Oracle, ReferenceGPU
    orc = Oracle()
    osp = orc.tiles(noa, noa, nva, nva, ts)
    sp = drv.setup_mo_space(noa, noa, nva, nva, ts)
    tasks, fac, _ = drv.enumerate_tasks(sp, True)
    n = len(tasks) if ntasks < 0 else min(ntasks, len(tasks))
    evl = orbital_energies(noa, noa, nva, nva)
    n_orb = [noa, noa, nva, nva]
    flops = np.array([task_flops(orc, osp, t) for t in tasks[:n]])

    # ours: one run_tasks call per task -> per-task kernel seconds and energies
    ctx = drv.Context(0)
    ctx.set_space(sp, evl, True)
    ctx.set_synthetic(SEED)
    ctx.run_tasks(np.arange(min(n, 2), dtype=np.int64))  # warm-up
    ours_ms, ours_e = [], []
    for i in range(n):
        e1, e2, st, _ = ctx.run_tasks(np.array([i], np.int64))
        ours_ms.append(st["seconds_kernel"] * 1e3)
        ours_e.append((e1, e2))
    ctx.close()
    ours_ms, ours_e = np.array(ours_ms), np.array(ours_e)

    res = {"shape": name, "nocc": noa, "nvir": nva, "ccsdt_tilesize": ts, "kernel_tasks": int(n),
           "of_tasks": int(len(tasks)), "counted_flops": float(flops.sum()),
           "ours": {"kernel": "fused_t_dmma_kernel", "kernel_ms": float(ours_ms.sum()),
                    "tflops": float(flops.sum() / ours_ms.sum() / 1e9)},
           "per_task": []}
    for kind in kinds:
        ref = ReferenceGPU(kind)
        out, nrec = ref.execute_synth(osp, evl, n_orb, SEED, True, tilesize=ts, task_limit=(-1 if ntasks < 0 else n))
        tr = ref.last_trace[:n]
        ref.release()
        ms = tr[:, 7]
        e_ref = tr[:, 8:10] * fac[:n, None]            # kernel partial sums x factor = task energies
        err = np.abs(e_ref - ours_e).max()
        rel = (np.abs(e_ref - ours_e) / np.maximum(1.0, np.abs(e_ref))).max()
        res[f"reference_{kind}"] = {
            "kernel": "fully_fused_kernel_ccsd_t_nvidia_tc_fp64 (K1, DMMA)" if kind == "tc" else
                      "revised_jk_ccsd_t_fully_fused_kernel (K2, FMA)",
            "kernel_ms": float(ms.sum()), "tflops": float(flops.sum() / ms.sum() / 1e9),
            "speedup_ours": float(ms.sum() / ours_ms.sum()),
            "max_abs_task_energy_diff_vs_ours": float(err), "max_rel_task_energy_diff_vs_ours": float(rel),
            "job_energies": None if np.isnan(out[0]) else [float(out[0]), float(out[1])]}
        for i in range(n):
            while len(res["per_task"]) <= i:
                j = len(res["per_task"])
                res["per_task"].append({"task": [int(x) for x in tasks[j][:6]],
                                        "extents": [int(x) for x in sp.k_range[tasks[j][:6]]],
                                        "flops": float(flops[j]), "ours_ms": float(ours_ms[j]),
                                        "ours_E": [float(x) for x in ours_e[j]]})
            res["per_task"][i][f"ref_{kind}_ms"] = float(ms[i])
            res["per_task"][i][f"ref_{kind}_E"] = [float(x) for x in e_ref[i]]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "comparator.json"))
    ap.add_argument("--skip-fma", action="store_true")
    ap.add_argument("--big-tasks", type=int, default=2)
    args = ap.parse_args()
    kinds = ["tc"] if args.skip_fma else ["tc", "fma"]
    results = [
        run_shape("small parity case (O=4,V=10 per spin, tile 6), whole job", 4, 10, 6, -1, kinds),
        run_shape("benzene cc-pVDZ shape (O=21,V=93 per spin, tile 40), whole job", 21, 93, 40, -1, kinds),
        run_shape(f"synthetic (60,500) tile 32, first {args.big_tasks} kernel tasks", 60, 500, 32, args.big_tasks, kinds),
        # BASELINE configs[2]: caffeine cc-pVDZ (O=51, V=195 per spin, ccsdt_tilesize 28 -> tiles 28,23 | 28x6,27)
        run_shape("caffeine cc-pVDZ shape (O=51,V=195 per spin, tile 28), first 6 kernel tasks", 51, 195, 28, 6, kinds),
    ]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    for r in results:
        line = f"{r['shape']}: ours {r['ours']['kernel_ms']:.2f} ms ({r['ours']['tflops']:.2f} TF/s)"
        for kind in kinds:
            k = r[f"reference_{kind}"]
            line += (f" | ref {kind} {k['kernel_ms']:.2f} ms ({k['tflops']:.2f} TF/s, ours x{k['speedup_ours']:.2f}, "
                     f"max |dE| {k['max_abs_task_energy_diff_vs_ours']:.2e})")
        print(line, flush=True)


if __name__ == "__main__":
    main()
