#!/usr/bin/env python
"""bench.py -- (T) wall time and FP64 TFLOP/s of the fused CCSD(T) triples path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

A "step" is one complete pass of the hot path over the workload:
  N = 1 : BASELINE.json configs[1] -- the benzene cc-pVDZ shape (O=21, V=93 per spin,
          ccsdt_tilesize 40 -> 28 kernel tasks, 1.59e13 counted flops), the WHOLE job, on synthetic
          spin-orbital amplitudes/integrals of that shape (no converged amplitudes exist offline).
  N > 1 : BASELINE.json configs[4] at the north_star target size -- the synthetic (nocc, nvir) = (60, 500)
          problem at ccsdt_tilesize 32 (16 320 kernel tasks, 1.24e17 counted flops in all): a strided sample
          of TASKS_PER_GPU*N kernel tasks of its canonical list (every (n_tasks/sample)-th task, so the
          sample has the job's mix of tile shapes and spin cases), handed out dynamically across the N
          ranks through a process-shared counter (weak scaling: work per GPU is fixed); the only
          collective is one NCCL all-reduce of the two energies at the end of the step.
`value`  = counted flops (the reference's own total_num_ops formula) / step time, tensors resident
           in HBM (N=1) or generated on the device (N>1) before the timed region.
`e2e`    = the same metric through CCSD_T_Fused_Driver.execute-style use of the C ABI with HOST
           tensors: H2D of all five tensors from pinned memory inside the timed region and the
           energies read back (N=1; at N>1 the 360 GB v2iabc of the workload cannot be host resident).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "(T) FP64 TFLOP/s (counted flops / (T) wall time)"
SEED = 1234
BENZENE = dict(name="benzene cc-pVDZ shape (BASELINE configs[1])", noa=21, nob=21, nva=93, nvb=93, ts=40)
SYNTH = dict(name="synthetic (nocc,nvir)=(60,500) (BASELINE configs[4], north_star target size)",
             noa=60, nob=60, nva=500, nvb=500, ts=32)
TASKS_PER_GPU = 24
CPU_SAMPLE_TS = 14
CPU_SAMPLE_TASKS = 3
# dram__bytes_read.sum + dram__bytes_write.sum of one fused-kernel launch (ncu --set full, profiles/): the
# N=1 workload's largest task, and the N>1 workload's task 5000
NCU_TRAFFIC = {"benzene": {"bytes": 5.87e8, "task": "task 0 (21,21,21,40,40,40), symmetry on: 9 975 of 166 375 boxes, 2.88 ms", "report": "profiles/ncu_r01_benzene_task0.txt"},
               "synth": {"bytes": 9.79e10, "task": "task 5000 (32,28,28,32,32,20), no coinciding tiles", "report": "profiles/ncu_r01_n60v500_task5000.txt"}}


def orbital_energies(w):
    from exachem_b200 import synthetic as syn
    return syn.Orbitals(w["noa"], w["nob"], w["nva"], w["nvb"]).orbital_energies()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w": float(np.median(power)) if power else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def task_flops(orc, osp, restricted, task):
    s1, d1, d2, _ = orc.task_exec(osp, restricted, task)
    base = 2.0 * float(np.prod(osp.k_range[task[:6]].astype(float)))
    f = base * int((s1 >= 0).sum())
    for h7 in range(osp.noab):
        f += base * int((d1[9 * h7:9 * h7 + 9] >= 0).sum()) * int(osp.k_range[h7])
    for p7 in range(osp.nvab):
        f += base * int((d2[9 * p7:9 * p7 + 9] >= 0).sum()) * int(osp.k_range[osp.noab + p7])
    return f


def cpu_reference_sample(w, ntasks=CPU_SAMPLE_TASKS, ts=CPU_SAMPLE_TS):
    """Times the reference's own CPU path (oracle/_ref, else the oracle port) on the first `ntasks`
    kernel tasks of workload `w` re-tiled at tile size `ts` (bounded sample: the CPU kernel needs two
    T^6 buffers per task and ~2e10 flop/s on all cores, so full-size tiles are out of reach)."""
    from oracle.oracle import Oracle
    orc = Oracle()
    osp = orc.tiles(w["noa"], w["nob"], w["nva"], w["nvb"], ts)
    tasks, _, _ = orc.enumerate(osp, True)
    flops = sum(task_flops(orc, osp, True, t) for t in tasks[:ntasks])
    evl = orbital_energies(w)
    n_orb = [w["noa"], w["nob"], w["nva"], w["nvb"]]
    try:
        from oracle.oracle import Reference
        ref = Reference()
        kind, cores = "reference", ref.num_threads()
        t0 = time.perf_counter()
        ref.execute_synth(osp, evl, n_orb, SEED, True, tilesize=ts, task_limit=ntasks)
        dt = time.perf_counter() - t0
    except (FileNotFoundError, OSError):
        from exachem_b200 import synthetic as syn
        kind, cores = "port", os.cpu_count() or 1
        T = syn.dense_all(syn.Orbitals(*n_orb), SEED)
        t0 = time.perf_counter()
        orc.run(osp, T, True, 0, ntasks)
        dt = time.perf_counter() - t0
    return {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind, "seconds": dt,
            "sample": f"first {ntasks} kernel task(s) of the {w['name']} re-tiled at ccsdt_tilesize {ts} "
                      f"({flops:.3e} counted flops), the reference's CPU (T) kernel (total_fused_ccsd_t_cpu, OpenMP) "
                      f"on {cores} threads"}


def run_reference_arm(args, rank):
    w = BENZENE if args.gpus == 1 else SYNTH
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference CPU kernel is OpenMP and gets every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(w)
        if i >= args.warmup:
            times.append(last["seconds"])
    flops = last["value"] * 1e12 * last["seconds"]
    dt = float(np.mean(times))
    v = flops / dt / 1e12
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "nocc": w["noa"], "nvir": w["nva"], "ccsdt_tilesize": w["ts"],
                       "bounded_sample": last["sample"]},
            "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub", default="", help="CTA box override, e.g. 1,1,2")
    ap.add_argument("--workload", default="", choices=["", "benzene", "synth"], help="override the per-N default")
    ap.add_argument("--static", action="store_true", help="N>1: static cost-balanced split instead of the shared counter")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist
    from exachem_b200 import _lib, driver as drv, multigpu

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()

    wl = args.workload or ("benzene" if world == 1 else "synth")
    w = BENZENE if wl == "benzene" else SYNTH
    sp = drv.setup_mo_space(w["noa"], w["nob"], w["nva"], w["nvb"], w["ts"])
    evl = orbital_energies(w)
    tasks, _, _ = drv.enumerate_tasks(sp, True)
    if wl == "benzene":
        task_ids = np.arange(len(tasks), dtype=np.int64)                     # the whole job
    else:
        n_sample = min(len(tasks), TASKS_PER_GPU * world)
        task_ids = (np.arange(n_sample, dtype=np.int64) * len(tasks)) // n_sample   # strided sample

    opts = {"rank": rank, "nranks": world}
    if args.sub:
        opts["sub"] = tuple(int(x) for x in args.sub.split(","))

    # FP64 peaks of this GPU (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)
    peaks = {}
    tf, ms = C.c_double(0), C.c_double(0)
    for name, flag in (("dfma", 0), ("dmma", 1)):
        L.ccsdt_probe_fp64_peak(local, flag, 20000, C.byref(tf), C.byref(ms))
        peaks[name] = tf.value

    ctx = drv.Context(local)
    ctx.set_options(**opts)
    ctx.set_space(sp, evl, True)

    host = None
    n_orb = np.array([w["noa"], w["nob"], w["nva"], w["nvb"]])
    if wl == "benzene" and world == 1:
        # host tensors in pinned memory (produced by the device generator, read back once)
        O, V = int(sp.k_range[:sp.noab].sum()), int(sp.k_range[sp.noab:].sum())
        dims = {drv.T1: (V, O, 1, 1), drv.T2: (V, V, O, O), drv.V_IJAB: (O, O, V, V), drv.V_IJKA: (O, O, O, V),
                drv.V_IABC: (O, V, V, V)}
        host = {}
        for tid, d in dims.items():
            n = int(np.prod(d))
            buf = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
            lo = np.zeros(4, np.int64)
            nn = np.array(d, np.int64)
            rc = L.ccsdt_synth_block(local, SEED, tid, *[int(x) for x in n_orb], lo.ctypes.data_as(_lib._i64p),
                                     nn.ctypes.data_as(_lib._i64p), buf.ctypes.data_as(_lib._dp))
            assert rc == 0
            host[tid] = buf
        for tid, buf in host.items():
            ctx.put_dense(tid, buf)
    else:
        ctx.set_synthetic(SEED)

    counter = None
    if world > 1 and not args.static:
        name = f"ccsdt_b200_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            counter = multigpu.SharedTaskCounter(name, create=True)
        dist.barrier()
        if rank != 0:
            counter = multigpu.SharedTaskCounter(name, create=False)
        ctx.set_task_counter(counter.address)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def set_symmetry(on):
        ctx.set_options(**dict(opts, symmetry=1 if on else 0))

    def step_resident():
        if counter is not None:
            dist.barrier()
            if rank == 0:
                counter.reset()
            dist.barrier()
        e1, e2, st, _ = ctx.run_tasks(task_ids)
        e1, e2 = multigpu.combine_energies(e1, e2, device="cuda")   # the one collective: E[T], E(T)
        return e1, e2, st

    def step_e2e():
        for tid, buf in host.items():
            ctx.put_dense(tid, buf, async_=True)   # H2D from pinned host memory, inside the timed region; the all-alpha
        return ctx.run_tasks(task_ids)[:3]         # tasks start while the other spin blocks are still on the bus

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        agg = {"seconds_kernel": 0.0, "seconds_staging": 0.0, "counted_flops": 0.0, "evaluated_flops": 0.0,
               "kernel_launches": 0, "h2d_bytes": 0, "d2h_bytes": 0, "tasks_run": 0}
        e = None
        for _ in range(steps):
            e1, e2, st = fn()
            e = (e1, e2)
            for k in agg:
                agg[k] += st[k]
        ev1.record()
        barrier()
        dt = time.perf_counter() - t0
        agg["device_ms"] = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            f = torch.tensor([agg["counted_flops"], agg["kernel_launches"], agg["tasks_run"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(f)
            agg["flops_all"], agg["launches_all"], agg["tasks_all"] = float(f[0].item()), int(f[1].item()), int(f[2].item())
        else:
            agg["flops_all"], agg["launches_all"], agg["tasks_all"] = agg["counted_flops"], agg["kernel_launches"], agg["tasks_run"]
        return dt, agg, e

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dt, agg, energies = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    # the same step with the symmetry reduction switched off (every CTA box evaluated, as the reference does):
    # the number that is comparable with the reference's kernels box for box
    n_off = max(1, min(args.steps, 2))
    set_symmetry(False)
    dt_off, agg_off, energies_off = timed(step_resident, n_off, 1)
    set_symmetry(True)

    if host is not None:
        n2 = max(1, min(args.steps, 3))
        dt2, agg2, _ = timed(step_e2e, n2, 1)
        e2e = {"value": agg2["flops_all"] / dt2 / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(agg2["h2d_bytes"] / n2),
               "host_tensor_bytes": int(sum(b.nbytes for b in host.values())),
               "d2h_bytes_per_step": int(agg2["d2h_bytes"] / n2), "ms_per_step": dt2 / n2 * 1e3,
               "api": "Context.put_dense(async) x5 (pinned dense host tensors; only their spin-conserving blocks, the ones the path "
                      "reads, cross the bus) + Context.run_tasks, i.e. what CCSD_T_Fused_Driver.execute does with dense host tensors"}
    else:
        e2e = {"value": agg["flops_all"] / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(agg["d2h_bytes"] / args.steps),
               "note": "tensors are generated on the device (v2iabc of this workload is 360 GB and cannot be "
                       "host resident); the host-tensor end-to-end number is measured at N=1"}

    if rank == 0:
        value = agg["flops_all"] / dt / 1e12
        peak = peaks["dmma"]
        kernel_tf = agg["evaluated_flops"] / max(agg["seconds_kernel"], 1e-12) / 1e12
        kernel_tf_off = agg_off["counted_flops"] / max(agg_off["seconds_kernel"], 1e-12) / 1e12
        traffic = NCU_TRAFFIC[wl]
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "nocc": w["noa"], "nvir": w["nva"], "ccsdt_tilesize": w["ts"],
                       "kernel_tasks_per_step": int(len(task_ids)),
                       "task_selection": "whole job" if wl == "benzene" else
                       f"strided sample: every {len(tasks) // len(task_ids)}-th of {len(tasks)} kernel tasks",
                       "parallelism": f"task-parallel x{world}" + ("" if world == 1 else
                                                                   (", static LPT split" if args.static else ", shared-counter dynamic hand-out")),
                       "l2_policy": "inputs larger than L2: the operand panels of one task (0.4-2.7 GB) exceed the 126 MB L2 "
                                    "and are rebuilt per task",
                       "cta_box": args.sub or "default (2,2,2,8,8,8), 3 CTAs/SM",
                       "series_note": "by the bench contract the N=1 line is BASELINE configs[1] (benzene shape) and the N>1 lines are "
                                      "configs[4] ((60,500), the north_star target): different workloads; the single-GPU point of the "
                                      "N>1 workload is `bench.py --workload synth` (profiles/bench_r01_n1_synth60x500.json, 56.7 TFLOP/s)"},
            "t_wall_s_per_step": dt / args.steps,
            "fraction_of_fp64_peak": value / (world * peak),
            "symmetry": {
                "what": "when two hole (particle) tiles of a task coincide the summand is symmetric under exchange of the "
                        "two indices, so the fused kernel evaluates only CTA boxes with ascending box coordinates and "
                        "weights them (options.symmetry, default on); `value` counts flops as the reference does "
                        "(every element), so it can exceed the FP64 peak",
                "evaluated_fraction": agg["evaluated_flops"] / max(agg["counted_flops"], 1.0),
                "value_symmetry_off": agg_off["flops_all"] / dt_off / 1e12,
                "ms_per_step_symmetry_off": dt_off / n_off * 1e3,
                "fraction_of_fp64_peak_symmetry_off": agg_off["flops_all"] / dt_off / 1e12 / (world * peak),
                "kernel_tflops_symmetry_off": kernel_tf_off,
                "max_abs_energy_diff_on_vs_off": max(abs(energies[0] - energies_off[0]), abs(energies[1] - energies_off[1])),
                "max_rel_energy_diff_on_vs_off": max(abs(energies[0] - energies_off[0]) / abs(energies_off[0]),
                                                     abs(energies[1] - energies_off[1]) / abs(energies_off[1]))},
            "energies": {"E[T]": energies[0], "E(T)": energies[1]},
            "roofline": {"bound": "tensor", "achieved": kernel_tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": kernel_tf / peak, "traffic": traffic["bytes"],
                         "traffic_note": f"DRAM read+write bytes of ONE fused-kernel launch ({traffic['task']}) from {traffic['report']}",
                         "kernel": "fused_t_dmma_kernel (FP64 DMMA m8n8k4, sm_100a)",
                         "peak_source": "measured in this run by ccsdt_probe_fp64_peak: register-resident DMMA.8x8x4 issue loop "
                                        "(MEASURED_PEAKS.json has no FP64 entry; nominal HGX B200 FP64 = 37 TFLOP/s)",
                         "peaks_measured": peaks,
                         "achieved_def": "reference-counted flops of the t3 elements this rank's fused-kernel launches had to "
                                         "evaluate (counted flops x CTA boxes evaluated / CTA boxes of the tile; padding not "
                                         "counted) / CUDA-event time of those launches on their stream (rank 0)",
                         "achieved_symmetry_off": kernel_tf_off, "frac_symmetry_off": kernel_tf_off / peak},
            "e2e": e2e, "gpu_launches": int(agg["launches_all"]), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_sample(w)
                cb.pop("seconds", None)
                line["cpu_baseline"] = cb
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "error": repr(ex)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if counter is not None:
        dist.barrier()
        counter.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
