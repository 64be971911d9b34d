#!/usr/bin/env python
"""bench.py -- (T) wall time and FP64 TFLOP/s of the fused CCSD(T) triples path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

A "step" is one complete pass of the hot path over the workload.  ONE workload at every N (strong scaling):

  BASELINE.json configs[4] at the north_star target size -- the synthetic (nocc, nvir) = (60, 500) problem at
  ccsdt_tilesize 32 (16 320 kernel tasks, 1.24e17 counted flops in all): a FIXED sample of its task list, all kernel
  tasks of 18 particle-tile triples (p4b,p5b,p6b) taken at equal strides through the triple list (about 100 tasks,
  the job's mix of tile shapes, spin cases and diagonal tiles; tasks of one triple share their v2iabc / T2 blocks, as in
  the whole job).  The same tasks run at N = 1, 2, 4, 8; across ranks they are handed out dynamically through a
  process-shared counter, and the only collective is one ncclAllReduce of the two energies (inside the C ABI).

`value`   = counted flops (the reference's own total_num_ops formula) of the sample / step time, tensors generated on
            the device before the timed region.
`e2e`     = the same metric through the reference-facing plug-in: CCSD_T_Fused_Driver<double>::execute of the C++
            drop-in header (include/ccsd_t_fused_driver_b200.hpp, compiled against the TAMM stand-in in tests/cpp), with
            HOST tensors: every block the tasks touch is pulled through Tensor::get -> fetch callback -> pinned ring ->
            H2D inside the timed region, a fresh context (cold HBM block store) per step, the energies read back.
            Across ranks the header's single-node defaults apply: shared-counter hand-out and the node-shared block
            store (ccsdt_share_attach: a block crosses PCIe once per node, the other ranks copy it from the fetching
            rank's HBM over NVLink); h2d_bytes_per_step and peer_bytes_per_step are summed over the ranks.
`benzene` = (N = 1 only) BASELINE.json configs[1], the benzene cc-pVDZ shape (O=21, V=93 per spin, ccsdt_tilesize 40 ->
            28 kernel tasks, 1.59e13 counted flops), the WHOLE job: its own value / e2e (through the same C++ header, dense
            host tensors) / roofline, as a second, named line inside the JSON.
Other workloads (--workload caffeine | gc | sweep:<nocc>,<nvir>,<ts> [--tasks M]) print the same line for the named
configs of BASELINE.json (profiles/ holds the committed runs).
"""
from __future__ import annotations

import argparse
import ctypes as C
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
BENZENE = dict(key="benzene", name="benzene cc-pVDZ shape (BASELINE configs[1])", noa=21, nob=21, nva=93, nvb=93, ts=40)
SYNTH = dict(key="synth", name="synthetic (nocc,nvir)=(60,500) (BASELINE configs[4], north_star target size)",
             noa=60, nob=60, nva=500, nvb=500, ts=32)
CAFFEINE = dict(key="caffeine", name="caffeine cc-pVDZ shape (BASELINE configs[2]; inputs/caffeine.json: ccsdt_tilesize 28)",
                noa=51, nob=51, nva=195, nvb=195, ts=28)
GC = dict(key="gc", name="guanine-cytosine base pair shape (BASELINE configs[3]; 6-31++G*, O=68, V=304 estimated, ts 40)",
          noa=68, nob=68, nva=304, nvb=304, ts=40)
N_TRIPLES = 18
CPU_SAMPLE_TS = 14
CPU_SAMPLE_TASKS = 3
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused-kernel launch from an `ncu --set full` capture (profiles/)
NCU_TRAFFIC = {"benzene": {"bytes": 2.639e9, "task": "task 8 (21,21,21 | 40,40,40 alpha-alpha-beta), the longest task of the job, 23.2 ms", "report": "profiles/ncu_r02_benzene_task8.txt"},
               "synth": {"bytes": 9.888e10, "task": "task 5000 (32,28,28,32,32,20), no coinciding tiles, 164.3 ms", "report": "profiles/ncu_r02_n60v500_task5000.txt"},
               "caffeine": {"bytes": 2.911e10, "task": "the largest task with six different execution tiles (26,26,25 | 40,40,40), 137.2 ms", "report": "profiles/ncu_r02_caffeine_exec40.txt"}}


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


def first_kernel_tasks(osp, n, restricted=True):
    """the first n kernel tasks in the reference's loop order (ccsd_t_fused_driver.hpp:368-395) without building the list
    (millions of entries for the small re-tiled tiles of the CPU sample)"""
    out, noab, pend, ks = [], osp.noab, osp.noab + osp.nvab, osp.k_spin
    for h1 in range(noab):
        for p4 in range(noab, pend):
            for h2 in range(h1, noab):
                for p5 in range(p4, pend):
                    for p6 in range(p5, pend):
                        for h3 in range(h2, noab):
                            sh, spn = ks[h1] + ks[h2] + ks[h3], ks[p4] + ks[p5] + ks[p6]
                            if sh == spn and (not restricted or sh + spn <= 8):
                                out.append(np.array([h1, h2, h3, p4, p5, p6, 0], np.int64))
                                if len(out) == n:
                                    return out
    return out


def cpu_reference_sample(w):
    """Times the reference's own CPU path (oracle/_ref, else the oracle port) on the first kernel tasks of workload `w`
    re-tiled at a small tile size (bounded sample of a few seconds: the CPU kernel needs two T^6 buffers per task and runs
    at ~1e10 flop/s on all cores, so full-size tiles are out of reach)."""
    from oracle.oracle import Oracle
    ts, ntasks = (10, 4) if w["nva"] >= 300 else (CPU_SAMPLE_TS, CPU_SAMPLE_TASKS)
    orc = Oracle()
    osp = orc.tiles(w["noa"], w["nob"], w["nva"], w["nvb"], ts)
    flops = sum(task_flops(orc, osp, True, t) for t in first_kernel_tasks(osp, ntasks))
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


def workload_config(w, **extra):
    return dict({"workload": w["name"], "nocc": w["noa"], "nvir": w["nva"], "ccsdt_tilesize": w["ts"]}, **extra)


def run_reference_arm(args, rank, w):
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, bounded_sample=last["sample"]),
            "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def select_tasks(tasks, spec, world):
    """task ids of the step.  spec: 'all' | ('triples', n) | ('strided', m)"""
    if spec == "all":
        return np.arange(len(tasks), dtype=np.int64), "whole job"
    kind, n = spec
    if kind == "strided":
        m = min(len(tasks), n)
        return (np.arange(m, dtype=np.int64) * len(tasks)) // m, f"strided sample: every {len(tasks) // m}-th of {len(tasks)} kernel tasks"
    triples = np.unique(tasks[:, 3:6], axis=0)                   # sorted (p4b, p5b, p6b)
    n = min(n, len(triples))
    pick = triples[(np.arange(n) * len(triples)) // n]
    keep = np.zeros(len(tasks), bool)
    for t in pick:
        keep |= (tasks[:, 3:6] == t).all(axis=1)
    ids = np.nonzero(keep)[0].astype(np.int64)
    return ids, (f"fixed sample, the same at every N: all {len(ids)} kernel tasks of {n} particle-tile triples taken at equal "
                 f"strides through the {len(triples)} triples of the job ({len(tasks)} kernel tasks)")


class Harness:
    """the C++ drop-in header behind a C entry (tests/cpp/adapter_harness.cpp)"""

    def __init__(self):
        from exachem_b200 import _lib
        so = os.path.join(ROOT, "tests", "cpp", "_build", "libadapter_test.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
        _lib.load()
        L = C.CDLL(so)
        dp, i64p, i32p = _lib._dp, _lib._i64p, _lib._i32p
        L.adapter_last_error.restype = C.c_char_p
        L.adapter_ccsdt_execute_ex.restype = C.c_int
        L.adapter_ccsdt_execute_ex.argtypes = ([C.c_int] * 4 + [i64p, i32p] + [dp] * 6 + [C.c_int, C.c_int, dp, i64p, C.POINTER(C.c_longdouble),
                                               C.POINTER(_lib.Stats), i64p, C.c_int64, C.POINTER(_lib.Options)])
        L.adapter_table_open.restype = C.c_void_p
        L.adapter_table_open.argtypes = [C.c_int] * 4 + [i64p, i32p, dp, C.c_uint64, C.c_int]
        L.adapter_table_close.argtypes = [C.c_void_p]
        L.adapter_table_info.argtypes = [C.c_void_p, i64p]
        L.adapter_table_execute.restype = C.c_int
        L.adapter_table_execute.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i64p, C.c_int64, C.POINTER(_lib.Options), dp,
                                            C.POINTER(_lib.Stats), i64p]
        self.L, self._lib = L, _lib

    def options(self, **kw):
        o = self._lib.Options()
        self._lib.load().ccsdt_default_options(C.byref(o))
        o.exec_tilesize = -1
        for k, v in kw.items():
            if k == "sub":
                o.sub[0], o.sub[1], o.sub[2] = v
            else:
                setattr(o, k, v)
        return o


def stats_dict(st, _lib):
    return {k: getattr(st, k) for k, _ in _lib.Stats._fields_}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-benzene", action="store_true", help="N=1: skip the second, named benzene line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sub", default="", help="CTA box override, e.g. 1,1,2")
    ap.add_argument("--workload", default="synth", help="synth (default) | benzene | caffeine | gc | sweep:<nocc>,<nvir>,<ts>")
    ap.add_argument("--tasks", type=int, default=0, help="strided sample of this many kernel tasks instead of the workload's default")
    ap.add_argument("--exec-tilesize", type=int, default=None, help="ccsdt_options.exec_tilesize (default: -1 = auto)")
    ap.add_argument("--static", action="store_true", help="N>1: static cost-balanced split instead of the shared counter")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload.startswith("sweep:"):
        no, nv, ts = (int(x) for x in args.workload[6:].split(","))
        w = dict(key="sweep", name=f"synthetic (nocc,nvir)=({no},{nv}) ccsdt_tilesize {ts} (BASELINE configs[4] sweep)", noa=no, nob=no,
                 nva=nv, nvb=nv, ts=ts)
        spec = ("strided", args.tasks or 48)
    else:
        w = {"synth": SYNTH, "benzene": BENZENE, "caffeine": CAFFEINE, "gc": GC}[args.workload]
        spec = ("triples", N_TRIPLES) if w is SYNTH else "all"
        if args.tasks:
            spec = ("strided", args.tasks)
    if args.impl == "reference":
        run_reference_arm(args, rank, w)
        return

    # stdout carries ONE JSON line: libraries that write there on their own (NCCL announces its version) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from exachem_b200 import _lib, driver as drv, multigpu

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()
    H = Harness()
    exec_ts = -1 if args.exec_tilesize is None else args.exec_tilesize

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 peaks of this GPU (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)
    peaks = {}
    tf, ms = C.c_double(0), C.c_double(0)
    for name, flag in (("dfma", 0), ("dmma", 1)):
        L.ccsdt_probe_fp64_peak(local, flag, 20000, C.byref(tf), C.byref(ms))
        peaks[name] = tf.value
    peak = peaks["dmma"]

    base_opts = {"rank": rank, "nranks": world, "exec_tilesize": exec_ts}
    if args.sub:
        base_opts["sub"] = tuple(int(x) for x in args.sub.split(","))

    def timed(fn, steps, warmup):
        """W warm-up steps, then K steps between barrier + synchronize on both sides; the max over ranks of the wall time"""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        agg, e = {}, None
        for _ in range(steps):
            e1, e2, st = fn()
            e = (e1, e2)
            for k, v in st.items():
                agg[k] = agg.get(k, 0) + v
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            keys = ["counted_flops", "kernel_launches", "tasks_run", "h2d_bytes", "d2h_bytes", "evaluated_flops", "executed_flops",
                    "blocks_from_peers", "peer_bytes"]
            f = torch.tensor([agg[k] for k in keys], dtype=torch.float64, device="cuda")
            dist.all_reduce(f)
            for k, v in zip(keys, f.tolist()):
                agg[k + "_all"] = v
        else:
            for k in list(agg):
                agg[k + "_all"] = agg[k]
        return dt, agg, e

    def measure(w, spec, steps, warmup, with_e2e, dense_host):
        """one workload: device-resident value, the same step with the symmetry reduction off, and the e2e through the
        C++ header.  dense_host: the e2e host tensors are dense arrays (benzene) instead of a block table."""
        sp = drv.setup_mo_space(w["noa"], w["nob"], w["nva"], w["nvb"], w["ts"])
        evl = orbital_energies(w)
        ctx = drv.Context(local)
        ctx.set_options(**base_opts)
        ctx.set_space(sp, evl, True)
        host = None
        if dense_host:
            # dense host tensors (produced once by the device generator): resident in HBM for `value`, pulled block by block
            # through Tensor::get for `e2e`
            O, V = int(sp.k_range[:sp.noab].sum()), int(sp.k_range[sp.noab:].sum())
            dims = {drv.T1: (V, O, 1, 1), drv.T2: (V, V, O, O), drv.V_IJAB: (O, O, V, V), drv.V_IJKA: (O, O, O, V),
                    drv.V_IABC: (O, V, V, V)}
            n_orb = [w["noa"], w["nob"], w["nva"], w["nvb"]]
            host = {}
            for tid, d in dims.items():
                buf = np.zeros(int(np.prod(d)))
                lo, nn = np.zeros(4, np.int64), np.array(d, np.int64)
                assert L.ccsdt_synth_block(local, SEED, tid, *n_orb, lo.ctypes.data_as(_lib._i64p),
                                           nn.ctypes.data_as(_lib._i64p), buf.ctypes.data_as(_lib._dp)) == 0
                host[tid] = buf
            for tid, buf in host.items():
                ctx.put_dense(tid, buf)
        else:
            ctx.set_synthetic(SEED)
        ex = ctx.exec_space()
        tasks, _, _ = drv.enumerate_tasks(ex, True)
        task_ids, selection = select_tasks(tasks, spec, world)
        retiled = not np.array_equal(ex.k_range, sp.k_range)
        # flops as the reference counts them on ITS tiling (whole job), or of the sampled tasks (execution tiling)
        ref_flops = float(drv.count_ops(sp, True)) if spec == "all" else None

        counter = None
        if world > 1 and not args.static:
            name = f"ccsdt_b200_{os.environ.get('MASTER_PORT', '0')}_{w['key']}"
            if rank == 0:
                counter = multigpu.SharedTaskCounter(name, create=True)
            dist.barrier()
            if rank != 0:
                counter = multigpu.SharedTaskCounter(name, create=False)
            ctx.set_task_counter(counter.address)
        if world > 1:
            # the one collective lives in the C ABI: ncclAllReduce of {E[T], E(T)} (ccsdt_comm_*)
            uid = C.create_string_buffer(128)
            if rank == 0:
                assert L.ccsdt_comm_unique_id(uid) == 0
            box = [uid.raw]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0], rank, world)

        def step_resident():
            if counter is not None:
                dist.barrier()
                if rank == 0:
                    counter.reset()
                dist.barrier()
            e1, e2, st, _ = ctx.run_tasks(task_ids)
            if world > 1:
                e1, e2 = ctx.comm_allreduce(e1, e2)
            return e1, e2, st

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        dt, agg, energies = timed(step_resident, steps, warmup)
        clocks = sampler.stop() if rank == 0 else None
        # the same step with the symmetry reduction switched off (every CTA box evaluated, as the reference does):
        # the number that is comparable with the reference's kernels box for box
        ctx.set_options(**dict(base_opts, symmetry=0))
        dt_off, agg_off, energies_off = timed(step_resident, 1, 1 if steps > 2 else 0)
        ctx.close()
        if counter is not None:
            dist.barrier()
            counter.close()

        flops_step = ref_flops if ref_flops is not None else agg["counted_flops_all"] / steps
        value = flops_step * steps / dt / 1e12
        kernel_tf = agg["evaluated_flops"] / max(agg["seconds_kernel"], 1e-12) / 1e12
        kernel_tf_off = agg_off["counted_flops"] / max(agg_off["seconds_kernel"], 1e-12) / 1e12
        out = {
            "value": value, "ms_per_step": dt / steps * 1e3,
            "config": workload_config(
                w, kernel_tasks_per_step=int(len(task_ids)), task_selection=selection,
                execution_tiling=("re-cut: " + str([int(x) for x in ex.k_range[:ex.noa]]) + " / " +
                                  str([int(x) for x in ex.k_range[ex.noab:ex.noab + ex.nva]]) + " (occ / virt alpha), blocks in the caller's tiling")
                if retiled else "the caller's tiles",
                counted_flops_per_step=flops_step,
                counted_flops_def="total_num_ops of the reference on its own tiling" if ref_flops is not None else
                "reference flop formula summed over the sampled kernel tasks",
                parallelism=f"task-parallel x{world}" + ("" if world == 1 else (", static LPT split" if args.static else
                                                                                 ", shared-counter dynamic hand-out, longest task first")),
                l2_policy="inputs larger than L2: the operand panels of one task (0.4-2.7 GB) exceed the 126 MB L2 and are rebuilt per task",
                cta_box=args.sub or "default (2,2,2,8,8,8), 3 CTAs/SM"),
            "fraction_of_fp64_peak_counted": value / (world * peak),
            "symmetry": {
                "what": "when two hole (particle) tiles of a task coincide the summand is symmetric under exchange of the two "
                        "indices, so the fused kernel evaluates only CTA boxes with ascending box coordinates and weights them "
                        "(options.symmetry, default on); `value` counts flops as the reference does (every element), so it can "
                        "exceed the FP64 peak",
                "evaluated_fraction": agg["evaluated_flops_all"] / max(agg["counted_flops_all"], 1.0),
                "value_symmetry_off": agg_off["counted_flops_all"] / dt_off / 1e12,
                "ms_per_step_symmetry_off": dt_off * 1e3,
                "fraction_of_fp64_peak_symmetry_off": agg_off["counted_flops_all"] / dt_off / 1e12 / (world * peak),
                "kernel_tflops_symmetry_off": kernel_tf_off,
                "max_rel_energy_diff_on_vs_off": max(abs(energies[0] - energies_off[0]) / abs(energies_off[0]),
                                                     abs(energies[1] - energies_off[1]) / abs(energies_off[1]))},
            "energies": {"E[T]": energies[0], "E(T)": energies[1]},
            "roofline": {"bound": "tensor", "achieved": kernel_tf, "peak": peak, "unit": "TFLOP/s", "frac": kernel_tf / peak,
                         "traffic": NCU_TRAFFIC.get(w["key"], {}).get("bytes"),
                         "traffic_note": (f"DRAM read+write bytes of ONE fused-kernel launch ({NCU_TRAFFIC[w['key']]['task']}) from the ncu "
                                          f"--set full capture {NCU_TRAFFIC[w['key']]['report']}") if w["key"] in NCU_TRAFFIC else None,
                         "kernel": "fused_t_dmma_kernel (FP64 DMMA m8n8k4, sm_100a)",
                         "peak_source": "measured in this run by ccsdt_probe_fp64_peak: register-resident DMMA.8x8x4 issue loop "
                                        "(MEASURED_PEAKS.json has no FP64 entry; nominal HGX B200 FP64 = 37 TFLOP/s)",
                         "peaks_measured": peaks,
                         "achieved_def": "reference-counted flops of the t3 elements this rank's fused-kernel launches had to evaluate "
                                         "(counted flops x CTA boxes evaluated / CTA boxes of the tile; padding not counted) / CUDA-event "
                                         "time of those launches on their streams (rank 0)",
                         "executed_over_evaluated": agg["executed_flops"] / max(agg["evaluated_flops"], 1.0),
                         "executed_note": "DMMA flops issued (ragged tiles padded to the CTA box, K to 4) over the flops counted for the "
                                          "same elements: the padding the kernel pays; frac x this = DMMA issue rate / peak",
                         "achieved_symmetry_off": kernel_tf_off, "frac_symmetry_off": kernel_tf_off / peak},
            "gpu_launches": int(agg["kernel_launches_all"]), "clocks": clocks,
        }

        if with_e2e:
            n2 = max(1, min(steps, 3))
            kr, ks = np.ascontiguousarray(sp.k_range, np.int64), np.ascontiguousarray(sp.k_spin, np.int32)
            out4, gets, st = np.zeros(4), np.zeros(5, np.int64), _lib.Stats()
            sub = np.ascontiguousarray(task_ids, np.int64) if spec != "all" else None
            opt = H.options(**{k: v for k, v in base_opts.items() if k not in ("rank", "nranks")})
            # ranks of the C++ header's ExecutionContext (TAMM stand-in): rank / size from the environment, meeting in a
            # POSIX shared-memory segment.  On one node the header's defaults apply: shared-counter hand-out and the
            # node-shared block store (a block crosses PCIe once per node, the other ranks copy it over NVLink).
            shim_env = {}
            if world > 1:
                key = f"bench{os.environ.get('MASTER_PORT', '0')}{w['key']}"
                if rank == 0:
                    with open(f"/dev/shm/tamm_shim_{key}", "wb") as f:
                        f.write(b"\0" * 4096)
                dist.barrier()
                shim_env = dict(TAMM_SHIM_RANK=str(rank), TAMM_SHIM_SIZE=str(world), TAMM_SHIM_KEY=key, CCSDT_B200_COUNTER_KEY=key)
                # with a dynamic hand-out any rank may be asked for any block: every rank's host table then holds all blocks
                # of the sample (filled by one single-rank pass before the timed steps).  Without the host memory for
                # that, the e2e step falls back to the static split, where a rank only ever touches its own blocks.
                avail = 0
                for ln in open("/proc/meminfo"):
                    if ln.startswith("MemAvailable"):
                        avail = int(ln.split()[1]) * 1024
                e2e_dynamic = dense_host or avail > world * 40e9
                if not e2e_dynamic:
                    shim_env.update(CCSDT_B200_DYNAMIC="0")
            if dense_host:
                host_bytes = int(sum(b.nbytes for b in host.values()))

                def step_e2e():
                    rc = H.L.adapter_ccsdt_execute_ex(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p),
                                                      evl.ctypes.data_as(_lib._dp), *[host[t].ctypes.data_as(_lib._dp) for t in range(5)], 1,
                                                      w["ts"], out4.ctypes.data_as(_lib._dp), gets.ctypes.data_as(_lib._i64p), None, C.byref(st),
                                                      sub.ctypes.data_as(_lib._i64p) if sub is not None else None,
                                                      len(sub) if sub is not None else 0, C.byref(opt))
                    assert rc == 0, H.L.adapter_last_error()
                    return float(out4[0]), float(out4[1]), stats_dict(st, _lib)
                host_kind = "dense host arrays; Tensor::get copies a block out row by row"
            else:
                table = H.L.adapter_table_open(sp.noa, sp.nob, sp.nva, sp.nvb, kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p),
                                               evl.ctypes.data_as(_lib._dp), SEED, local)
                cnt = np.array([sp.noa, sp.nob, sp.nva, sp.nvb], np.int32)

                def step_e2e():
                    rc = H.L.adapter_table_execute(table, cnt.ctypes.data_as(_lib._i32p), 1, w["ts"],
                                                   sub.ctypes.data_as(_lib._i64p) if sub is not None else None,
                                                   len(sub) if sub is not None else 0, C.byref(opt), out4.ctypes.data_as(_lib._dp),
                                                   C.byref(st), gets.ctypes.data_as(_lib._i64p))
                    assert rc == 0, H.L.adapter_last_error()
                    return float(out4[0]), float(out4[1]), stats_dict(st, _lib)
                host_kind = ("host block table: every block the tasks touch is one contiguous host buffer (a local TAMM block), "
                             "filled by the device generator on first touch during the warm-up step; Tensor::get is one memcpy")
            if world > 1 and not dense_host and e2e_dynamic:
                step_e2e()                      # single-rank pass over the whole sample: fills this rank's host table
            os.environ.update(shim_env)
            dt2, agg2, e_e2e = timed(step_e2e, n2, 1)
            for k in shim_env:
                os.environ.pop(k, None)
            if world > 1:
                dist.barrier()
                if rank == 0:
                    try:
                        os.unlink(f"/dev/shm/tamm_shim_{key}")
                    except OSError:
                        pass
            if not dense_host:
                info = np.zeros(3, np.int64)
                H.L.adapter_table_info(table, info.ctypes.data_as(_lib._i64p))
                host_bytes = int(info[0])
                H.L.adapter_table_close(table)
            flops2 = ref_flops * n2 if ref_flops is not None else agg2["counted_flops_all"]
            out["e2e"] = {"value": flops2 / dt2 / 1e12, "unit": "TFLOP/s", "ms_per_step": dt2 / n2 * 1e3,
                          "h2d_bytes_per_step": int(agg2["h2d_bytes_all"] / n2), "d2h_bytes_per_step": int(agg2["d2h_bytes_all"] / n2),
                          "host_tensor_bytes_rank0": host_bytes, "blocks_fetched_per_step_rank0": int(agg2["blocks_fetched"] / n2),
                          "seconds_in_Tensor_get_per_step_rank0": agg2["seconds_fetch"] / n2,
                          "seconds_kernel_per_step_rank0": agg2["seconds_kernel"] / n2,
                          "blocks_from_peers_per_step": int(agg2.get("blocks_from_peers_all", 0) / n2),
                          "peer_bytes_per_step": int(agg2.get("peer_bytes_all", 0) / n2),
                          "max_rel_energy_diff_vs_device_generated": max(abs(e_e2e[0] - energies[0]) / abs(energies[0]),
                                                                         abs(e_e2e[1] - energies[1]) / abs(energies[1])),
                          "api": "CCSD_T_Fused_Driver<double>::execute of include/ccsd_t_fused_driver_b200.hpp (C++ drop-in header over "
                                 "the C ABI), called as exachem/cc/ccsd_t/ccsd_t.cpp:253-256 does; context created and destroyed inside "
                                 "every call, HBM block store cold at the start of every step; " + host_kind +
                                 ("" if world == 1 else ("; shared-counter hand-out and node-shared block store (the header's defaults on one node)"
                                                         if e2e_dynamic else "; static cost-balanced split across the ranks, private block stores"))}
        return out

    benz = None
    if world == 1 and w is SYNTH and not args.no_benzene:
        # the second, named line first, while the process is fresh; then every cached device resource is released so that
        # the main workload starts from a clean device too
        benz = measure(BENZENE, "all", min(args.steps, 10), 3, not args.no_e2e, dense_host=True)
        benz.pop("clocks", None)
        L.ccsdt_release_cached()
    res = measure(w, spec, args.steps, args.warmup, not args.no_e2e, dense_host=(w is BENZENE))
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update({k: v for k, v in res.items() if k not in ("value", "ms_per_step")})
        line["t_wall_s_per_step"] = res["ms_per_step"] * 1e-3
        if "e2e" not in line:
            line["e2e"] = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "--no-e2e"}
        if benz is not None:
            line["benzene"] = dict({"metric": METRIC, "unit": "TFLOP/s", "what": "BASELINE.json configs[1], the whole job on one GPU "
                                    "(the N=1 line of round 1): value with tensors resident, e2e through the C++ header from dense host tensors"},
                                   **benz)
    if rank == 0:
        if not args.no_cpu_baseline:
            try:
                # a clean process: torchrun pins OMP_NUM_THREADS=1 and the OpenMP runtime of this one has read it already
                env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
                for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                    env.pop(k, None)
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                      "--workload", args.workload], env=env, capture_output=True, text=True, timeout=900)
                ref_line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
                line["cpu_baseline"] = ref_line["cpu_baseline"]
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "error": repr(ex)}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
