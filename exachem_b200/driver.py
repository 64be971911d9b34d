"""Python mirror of the reference operator interface for the (T) path.

`CCSD_T_Fused_Driver.execute` / `.calculate_performance_ops` follow
exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:73-87 (same names, argument meaning and return tuple);
the TAMM objects are replaced by plain containers:

  MO            -> TiledSpace  (MO("occ").num_tiles(), MO.input_tile_sizes(), k_spin)
  Tensor<T>     -> a dense numpy array over the spin-orbital ranges, or any object with
                   `.get(bid) -> ndarray` (what Tensor::get delivers: one unsorted row-major block)
  V2Tensors<T>  -> dict / object with v2ijab, v2ijka, v2iabc
  LRUCache args -> accepted and ignored (the HBM block store replaces them)

Every call goes through the C ABI (include/ccsdt_b200.h); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import FETCH_FN, Options, Stats

T1, T2, V_IJAB, V_IJKA, V_IABC = 0, 1, 2, 3, 4
KERNEL_DMMA, KERNEL_SIMPLE = 0, 1


class CcsdtError(RuntimeError):
    pass


def lib():
    return _lib.load()


def _p(a, t):
    return a.ctypes.data_as(t)


@dataclass
class TiledSpace:
    noa: int
    nob: int
    nva: int
    nvb: int
    k_range: np.ndarray
    k_spin: np.ndarray

    @property
    def noab(self):
        return self.noa + self.nob

    @property
    def nvab(self):
        return self.nva + self.nvb

    @property
    def k_offset(self):
        return np.concatenate([[0], np.cumsum(self.k_range)]).astype(np.int64)

    def num_tiles(self, name: str) -> int:  # MO(name).num_tiles()
        return {"occ": self.noab, "virt": self.nvab, "occ_alpha": self.noa, "occ_beta": self.nob,
                "virt_alpha": self.nva, "virt_beta": self.nvb}[name]

    def input_tile_sizes(self):
        return self.k_range.tolist()


def setup_mo_space(n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, ccsdt_tilesize) -> TiledSpace:
    """setupMOIS(ec, chem_env, triples=true) + k_spin (cholesky_2e.cpp:186-279, ccsd_t.cpp:245-249)."""
    L = lib()
    cap = 8 + (n_occ_alpha + n_occ_beta + n_vir_alpha + n_vir_beta) // max(int(ccsdt_tilesize), 1)
    kr = np.zeros(cap, np.int64)
    ks = np.zeros(cap, np.int32)
    cnt = np.zeros(4, np.int32)
    n = L.ccsdt_tiles(n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, ccsdt_tilesize,
                      _p(kr, _lib._i64p), _p(ks, _lib._i32p), _p(cnt, _lib._i32p), cap)
    if n < 0:
        raise CcsdtError("ccsdt_tiles failed")
    return TiledSpace(int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3]), kr[:n].copy(), ks[:n].copy())


def enumerate_tasks(sp: TiledSpace, is_restricted: bool):
    L = lib()
    ks = np.ascontiguousarray(sp.k_spin, np.int32)
    n_outer = C.c_int64(0)
    n = L.ccsdt_enumerate(sp.noab, sp.nvab, _p(ks, _lib._i32p), int(is_restricted), None, None, 0,
                          C.byref(n_outer))
    tasks = np.zeros((max(n, 1), 7), np.int64)
    fac = np.zeros(max(n, 1))
    L.ccsdt_enumerate(sp.noab, sp.nvab, _p(ks, _lib._i32p), int(is_restricted), _p(tasks, _lib._i64p),
                      _p(fac, _lib._dp), n, C.byref(n_outer))
    return tasks[:n], fac[:n], int(n_outer.value)


def task_terms(sp: TiledSpace, is_restricted: bool, task):
    L = lib()
    ks = np.ascontiguousarray(sp.k_spin, np.int32)
    kr = np.ascontiguousarray(sp.k_range, np.int64)
    t = np.ascontiguousarray(task[:6], np.int64)
    s1 = np.zeros(9, np.uint8)
    d1 = np.zeros(9 * sp.noab, np.uint8)
    d2 = np.zeros(9 * sp.nvab, np.uint8)
    L.ccsdt_task_terms(sp.noab, sp.nvab, _p(ks, _lib._i32p), _p(kr, _lib._i64p), int(is_restricted),
                       _p(t, _lib._i64p), _p(s1, _lib._u8p), _p(d1, _lib._u8p), _p(d2, _lib._u8p))
    return s1, d1, d2


def count_ops(sp: TiledSpace, is_restricted: bool) -> int:
    L = lib()
    ks = np.ascontiguousarray(sp.k_spin, np.int32)
    kr = np.ascontiguousarray(sp.k_range, np.int64)
    v = C.c_longdouble(0)
    L.ccsdt_count_ops(sp.noab, sp.nvab, _p(ks, _lib._i32p), _p(kr, _lib._i64p), int(is_restricted), C.byref(v))
    return int(v.value)


def partition(sp: TiledSpace, is_restricted: bool, nranks: int) -> np.ndarray:
    L = lib()
    ks = np.ascontiguousarray(sp.k_spin, np.int32)
    kr = np.ascontiguousarray(sp.k_range, np.int64)
    n = len(enumerate_tasks(sp, is_restricted)[0])
    own = np.zeros(max(n, 1), np.int32)
    L.ccsdt_partition(sp.noab, sp.nvab, _p(ks, _lib._i32p), _p(kr, _lib._i64p), int(is_restricted), nranks,
                      _p(own, _lib._i32p), n)
    return own[:n]


class Context:
    """Owns one ccsdt_ctx (one GPU)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.ccsdt_create(C.byref(self.h), device)
        if rc:
            raise CcsdtError(self.L.ccsdt_last_error(None).decode())
        self._keep = []

    def _ck(self, rc):
        if rc:
            raise CcsdtError(self.L.ccsdt_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.ccsdt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        o = Options()
        self.L.ccsdt_default_options(C.byref(o))
        for k, v in kw.items():
            if k == "sub":
                o.sub[0], o.sub[1], o.sub[2] = v
            else:
                setattr(o, k, v)
        self._ck(self.L.ccsdt_set_options(self.h, C.byref(o)))

    def set_space(self, sp: TiledSpace, evl, is_restricted: bool):
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        ev = np.ascontiguousarray(evl, np.float64)
        self._ck(self.L.ccsdt_set_space(self.h, sp.noa, sp.nob, sp.nva, sp.nvb, _p(kr, _lib._i64p),
                                        _p(ks, _lib._i32p), _p(ev, _lib._dp), int(is_restricted)))

    def put_dense(self, tensor: int, arr, async_: bool = False):
        """async_: return without waiting for the copy; `arr` (pinned memory for a truly asynchronous copy) must stay
        unchanged until the next run returns (a reference is kept here until then)."""
        a = np.ascontiguousarray(arr, np.float64)
        if async_:
            self._pending_uploads = getattr(self, "_pending_uploads", []) + [a]
            self._ck(self.L.ccsdt_put_dense_async(self.h, tensor, _p(a, _lib._dp)))
        else:
            self._ck(self.L.ccsdt_put_dense(self.h, tensor, _p(a, _lib._dp)))

    def put_block(self, tensor: int, bid, arr):
        a = np.ascontiguousarray(arr, np.float64)
        b = np.zeros(4, np.uint32)
        b[:len(bid)] = bid
        self._ck(self.L.ccsdt_put_block(self.h, tensor, _p(b, _lib._u32p), _p(a, _lib._dp)))

    def set_fetch(self, fn):
        """fn(tensor, bid tuple) -> ndarray block (row-major)."""
        if fn is None:
            self._ck(self.L.ccsdt_set_fetch(self.h, FETCH_FN(0), None))
            return

        def _cb(_user, tensor, bid, dst, n):
            try:
                nd = 2 if tensor == T1 else 4
                blk = np.ascontiguousarray(fn(tensor, tuple(int(bid[i]) for i in range(nd))), np.float64)
                if blk.size != n:
                    return 2
                C.memmove(dst, blk.ctypes.data, n * 8)
                return 0
            except Exception:  # noqa: BLE001 - reported through the C status code
                return 1

        cb = FETCH_FN(_cb)
        self._keep.append(cb)
        self._ck(self.L.ccsdt_set_fetch(self.h, cb, None))

    def put_cholesky(self, chol):
        """setupV2Tensors on the device: chol[N, N, ncv] over all spin orbitals in tile order (cholVpr)."""
        a = np.ascontiguousarray(chol, np.float64)
        assert a.ndim == 3 and a.shape[0] == a.shape[1]
        self._ck(self.L.ccsdt_put_cholesky(self.h, _p(a, _lib._dp), a.shape[2]))

    def set_synthetic(self, seed: int):
        self._ck(self.L.ccsdt_set_synthetic(self.h, seed))

    def clear_blocks(self):
        """drop every fetched block from the HBM block store: the next run starts cold, like a fresh execute"""
        self._ck(self.L.ccsdt_clear_blocks(self.h))

    def exec_space(self) -> TiledSpace:
        """the execution tiling in force (options exec_tilesize); task ids of run / run_tasks index ITS task list"""
        cap = 4096
        kr, ks, cnt = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(4, np.int32)
        n = self.L.ccsdt_exec_tiles(self.h, _p(kr, _lib._i64p), _p(ks, _lib._i32p), _p(cnt, _lib._i32p), cap)
        if n <= 0:
            raise CcsdtError("ccsdt_exec_tiles failed (no space set?)")
        return TiledSpace(int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3]), kr[:n].copy(), ks[:n].copy())

    def num_tasks(self) -> int:
        return int(self.L.ccsdt_num_tasks(self.h))

    def share_attach(self, name: str, local_rank: int, local_ranks: int, create: bool):
        """node-shared block store: blocks fetched by one rank of the node are read from its HBM by the others"""
        self._ck(self.L.ccsdt_share_attach(self.h, name.encode(), local_rank, local_ranks, int(create)))

    def share_detach(self):
        self._ck(self.L.ccsdt_share_detach(self.h))

    # ---- the one collective, inside the C ABI (NCCL) ----
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.L.ccsdt_comm_init(self.h, buf, rank, nranks))

    def comm_allreduce(self, e1: float, e2: float):
        e = np.array([e1, e2], np.float64)
        self._ck(self.L.ccsdt_comm_allreduce(self.h, _p(e, _lib._dp)))
        return float(e[0]), float(e[1])

    def set_task_counter(self, address: int | None):
        """address of a process-shared int64 (see multigpu.SharedTaskCounter); None = static split."""
        self._ck(self.L.ccsdt_set_task_counter(self.h, C.c_void_p(address) if address else None))

    def run(self, task_begin=0, task_end=-1, per_task_n=0):
        e = np.zeros(2)
        st = Stats()
        pt = np.zeros((per_task_n, 2)) if per_task_n else None
        self._ck(self.L.ccsdt_run(self.h, task_begin, task_end, _p(e, _lib._dp),
                                  _p(pt, _lib._dp) if pt is not None else None, C.byref(st)))
        stats = {k: getattr(st, k) for k, _ in Stats._fields_}
        self._pending_uploads = []
        return float(e[0]), float(e[1]), stats, pt


    def run_tasks(self, task_ids, per_task=False):
        """explicit list of task ids (indices into the canonical list)."""
        ids = np.ascontiguousarray(task_ids, np.int64)
        e = np.zeros(2)
        st = Stats()
        pt = np.zeros((len(ids), 2)) if per_task else None
        self._ck(self.L.ccsdt_run_tasks(self.h, _p(ids, _lib._i64p), len(ids), _p(e, _lib._dp),
                                        _p(pt, _lib._dp) if pt is not None else None, C.byref(st)))
        stats = {k: getattr(st, k) for k, _ in Stats._fields_}
        self._pending_uploads = []
        return float(e[0]), float(e[1]), stats, pt


class _BlockSource:
    """Adapts 'dense ndarray' or 'object with .get(bid)' to the Context operand API."""

    @staticmethod
    def attach(ctx: Context, tensor: int, src, pending_fetch: dict):
        if isinstance(src, np.ndarray):
            ctx.put_dense(tensor, src)
        elif hasattr(src, "get"):
            pending_fetch[tensor] = src
        else:
            raise TypeError(f"tensor {tensor}: need a dense ndarray or an object with .get(bid)")


@dataclass
class CCSD_T_Fused_Driver:
    """Mirror of `template<typename T> class CCSD_T_Fused_Driver` (ccsd_t_fused_driver.hpp:60-87)."""
    device: int = 0
    options: dict = field(default_factory=dict)
    last_stats: dict = field(default_factory=dict)

    def execute(self, chem_env, ec, k_spin, MO: TiledSpace, d_t1, d_t2, d_v2, k_evl_sorted,
                hf_ccsd_energy=0.0, is_restricted=True, cache_s1t=None, cache_s1v=None, cache_d1t=None,
                cache_d1v=None, cache_d2t=None, cache_d2v=None, seq_h3b=False, tilesize_opt=True):
        """Returns (energy1 = E[T] partial, energy2 = E(T) partial, work seconds, total seconds), the
        reference's tuple (ccsd_t_fused_driver.hpp:545); the caller sums partials over ranks
        (ccsd_t.cpp:262-263).  `ec` may carry .rank/.nranks (one process per GPU)."""
        sp = TiledSpace(MO.noa, MO.nob, MO.nva, MO.nvb, np.asarray(MO.k_range), np.asarray(k_spin))
        ctx = Context(self.device)
        try:
            opts = dict(self.options)
            opts.setdefault("rank", getattr(ec, "rank", 0) if ec is not None else 0)
            opts.setdefault("nranks", getattr(ec, "nranks", 1) if ec is not None else 1)
            ctx.set_options(**opts)
            ctx.set_space(sp, k_evl_sorted, is_restricted)
            fetch = {}
            v2 = d_v2 if isinstance(d_v2, dict) else {k: getattr(d_v2, k) for k in ("v2ijab", "v2ijka", "v2iabc")}
            for tid, src in ((T1, d_t1), (T2, d_t2), (V_IJAB, v2["v2ijab"]), (V_IJKA, v2["v2ijka"]),
                             (V_IABC, v2["v2iabc"])):
                _BlockSource.attach(ctx, tid, src, fetch)
            if fetch:
                ctx.set_fetch(lambda tensor, bid: fetch[tensor].get(bid))
            e1, e2, stats, _ = ctx.run()
            self.last_stats = stats
            return e1, e2, stats["seconds_kernel"] + stats["seconds_staging"], stats["seconds_total"]
        finally:
            ctx.close()

    def calculate_performance_ops(self, chem_env, ec, k_spin, MO: TiledSpace, k_evl_sorted=None,
                                  hf_ccsd_energy=0.0, is_restricted=True, seq_h3b=False) -> int:
        sp = TiledSpace(MO.noa, MO.nob, MO.nva, MO.nvb, np.asarray(MO.k_range), np.asarray(k_spin))
        return count_ops(sp, is_restricted)
