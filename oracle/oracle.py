"""TEST INFRASTRUCTURE ONLY.  ctypes front-ends for

  * oracle/_build/libccsdt_oracle.so  -- the C restatement (oracle/ccsdt_oracle.c), class `Oracle`
  * oracle/_ref/libccsdt_ref.so       -- the reference's own (T) path compiled unmodified against the
                                         TAMM shim (oracle/ref_driver.cpp), class `Reference`
  * oracle/_ref/libccsdt_refgpu_{tc,fma}.so -- the reference's own GPU task function and kernels (K1 DMMA, K2 FMA),
                                         unmodified, sm_100a, class `ReferenceGPU`: second checker on the GPU box and
                                         the GPU-side comparator (tests/test_gpu_comparator.py, tools/gpu_comparator.py,
                                         tools/benzene_real.py)
and an independent closed-form numpy statement of the 27 equations (`closed_form_energy`,
SURVEY.md §8 a9/a10) used to cross-check both.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libccsdt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libccsdt_ref.so")
REF_GPU_SO = {"tc": os.path.join(HERE, "_ref", "libccsdt_refgpu_tc.so"),
              "fma": os.path.join(HERE, "_ref", "libccsdt_refgpu_fma.so")}

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "ccsdt_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "_build/libccsdt_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def _p(a, t):
    return a.ctypes.data_as(t)


@dataclass
class Space:
    """Tiled spin-orbital space: what MO.input_tile_sizes()/k_spin give the reference driver."""
    noa: int
    nob: int
    nva: int
    nvb: int
    k_range: np.ndarray  # int64, len noab+nvab
    k_spin: np.ndarray   # int32

    @property
    def noab(self):
        return self.noa + self.nob

    @property
    def nvab(self):
        return self.nva + self.nvb

    @property
    def k_offset(self):
        return np.concatenate([[0], np.cumsum(self.k_range)[:-1]]).astype(np.int64)

    @property
    def Ot(self):
        return int(self.k_range[:self.noab].sum())

    @property
    def Vt(self):
        return int(self.k_range[self.noab:].sum())


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(ORACLE_SO)
        L.orc_tiles.argtypes = [C.c_int64] * 5 + [_i64p, _i32p, _i32p]
        L.orc_tiles.restype = C.c_int
        L.orc_enumerate.argtypes = [C.c_int, C.c_int, _i32p, C.c_int, _i64p, _dp, C.c_int64, _i64p]
        L.orc_enumerate.restype = C.c_int64
        L.orc_task_exec.argtypes = [C.c_int, C.c_int, _i32p, _i64p, C.c_int, _i64p, _i32p, _i32p, _i32p,
                                    _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.orc_task_exec.restype = None
        L.orc_count_ops.argtypes = [C.c_int, C.c_int, _i32p, _i64p, C.c_int]
        L.orc_count_ops.restype = C.c_longdouble
        L.orc_run.argtypes = [C.c_int, C.c_int, _i64p, _i32p] + [_dp] * 6 + [C.c_int, C.c_int64, C.c_int64,
                                                                           _dp, _dp]
        L.orc_run.restype = C.c_int
        self.L = L

    def tiles(self, n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, tilesize) -> Space:
        cap = 4 + (n_occ_alpha + n_occ_beta + n_vir_alpha + n_vir_beta) // max(tilesize, 1) + 4
        kr = np.zeros(cap, np.int64)
        ks = np.zeros(cap, np.int32)
        cnt = np.zeros(4, np.int32)
        n = self.L.orc_tiles(n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, tilesize, _p(kr, _i64p),
                             _p(ks, _i32p), _p(cnt, _i32p))
        return Space(int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3]), kr[:n].copy(), ks[:n].copy())

    def enumerate(self, sp: Space, is_restricted: bool):
        n_outer = C.c_int64(0)
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        n = self.L.orc_enumerate(sp.noab, sp.nvab, _p(ks, _i32p), int(is_restricted), None, None, 0,
                                 C.byref(n_outer))
        tasks = np.zeros((max(n, 1), 7), np.int64)
        fac = np.zeros(max(n, 1), np.float64)
        self.L.orc_enumerate(sp.noab, sp.nvab, _p(ks, _i32p), int(is_restricted), _p(tasks, _i64p),
                             _p(fac, _dp), n, C.byref(n_outer))
        return tasks[:n], fac[:n], int(n_outer.value)

    def task_exec(self, sp: Space, is_restricted: bool, task):
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        t = np.ascontiguousarray(task[:6], np.int64)
        s1 = np.zeros(9, np.int32)
        d1 = np.zeros(9 * sp.noab, np.int32)
        d2 = np.zeros(9 * sp.nvab, np.int32)
        cnt = np.zeros(3, np.int32)
        self.L.orc_task_exec(sp.noab, sp.nvab, _p(ks, _i32p), _p(kr, _i64p), int(is_restricted),
                             _p(t, _i64p), _p(s1, _i32p), _p(d1, _i32p), _p(d2, _i32p), _p(cnt, _i32p),
                             None, None, None, None, None)
        return s1, d1, d2, cnt

    def count_ops(self, sp: Space, is_restricted: bool) -> int:
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        v = self.L.orc_count_ops(sp.noab, sp.nvab, _p(ks, _i32p), _p(kr, _i64p), int(is_restricted))
        return int(v)

    def run(self, sp: Space, tensors: dict, is_restricted: bool, task_begin=0, task_end=-1,
            per_task=False):
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        arrs = [np.ascontiguousarray(tensors[k], np.float64)
                for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")]
        e = np.zeros(2)
        pt = None
        if per_task:
            n = self.enumerate(sp, is_restricted)[0].shape[0]
            pt = np.zeros((max(n, 1), 2))
        self.L.orc_run(sp.noab, sp.nvab, _p(kr, _i64p), _p(ks, _i32p), *[_p(a, _dp) for a in arrs],
                       int(is_restricted), task_begin, task_end, _p(e, _dp),
                       _p(pt, _dp) if pt is not None else None)
        return (e[0], e[1], pt) if per_task else (e[0], e[1])


class Reference:
    """The reference's own code (oracle/_ref).  Raises FileNotFoundError when it was never built."""

    def __init__(self, so_path: str = REF_SO):
        build()
        if not os.path.exists(so_path):
            raise FileNotFoundError(so_path)
        L = C.CDLL(so_path)
        L.ref_ccsdt_execute.argtypes = [C.c_int] * 4 + [_i64p, _i32p] + [_dp] * 6 + \
            [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp, C.c_int64, _i64p]
        L.ref_ccsdt_execute.restype = C.c_int
        L.ref_ccsdt_count_ops.argtypes = [C.c_int] * 4 + [_i64p, _i32p, C.c_int, C.POINTER(C.c_longdouble)]
        L.ref_ccsdt_task_info.argtypes = [C.c_int] * 4 + [_i64p, _i32p, C.c_int] + [C.c_int] * 6 + [_i32p] * 4
        L.ref_ccsdt_num_threads.restype = C.c_int
        L.ref_ccsdt_execute_synth.argtypes = [C.c_int] * 4 + [_i64p, _i32p, _dp, _i64p, C.c_uint64, C.c_int,
                                                               C.c_int, C.c_int, C.c_int64, _dp, _i64p, _dp,
                                                               C.c_int64]
        L.ref_ccsdt_execute_synth.restype = C.c_int
        self.L = L

    def num_threads(self) -> int:
        return int(self.L.ref_ccsdt_num_threads())

    def execute(self, sp: Space, tensors: dict, is_restricted: bool, tilesize=40, cache_size=8,
                enumerate_only=False, task_limit=-1):
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        arrs = [np.ascontiguousarray(tensors[k], np.float64)
                for k in ("evl", "t1", "t2", "v2ijab", "v2ijka", "v2iabc")]
        out = np.zeros(4)
        cap = 1 << 20
        trace = np.zeros((cap, 10))
        n = C.c_int64(0)
        self.L.ref_ccsdt_execute(sp.noa, sp.nob, sp.nva, sp.nvb, _p(kr, _i64p), _p(ks, _i32p),
                                 *[_p(a, _dp) for a in arrs], int(is_restricted), tilesize, cache_size,
                                 1 if enumerate_only else 0, task_limit, _p(out, _dp), _p(trace, _dp), cap,
                                 C.byref(n))
        return out, trace[:min(n.value, cap)].copy()

    def execute_synth(self, sp: Space, evl, n_orb, seed: int, is_restricted: bool, tilesize=40,
                      cache_size=8, task_limit=-1):
        """reference `execute` on procedural tensors; computes only the first task_limit kernel tasks"""
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        ev = np.ascontiguousarray(evl, np.float64)
        no = np.ascontiguousarray(n_orb, np.int64)
        out = np.zeros(4)
        n = C.c_int64(0)
        cap = 1 << 16
        trace = np.zeros((cap, 10))
        self.L.ref_ccsdt_execute_synth(sp.noa, sp.nob, sp.nva, sp.nvb, _p(kr, _i64p), _p(ks, _i32p), _p(ev, _dp),
                                       _p(no, _i64p), seed, int(is_restricted), tilesize, cache_size, task_limit,
                                       _p(out, _dp), C.byref(n), _p(trace, _dp), cap)
        self.last_trace = trace[:min(n.value, cap)].copy()
        return out, int(n.value)

    def count_ops(self, sp: Space, is_restricted: bool) -> int:
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        v = C.c_longdouble(0)
        self.L.ref_ccsdt_count_ops(sp.noa, sp.nob, sp.nva, sp.nvb, _p(kr, _i64p), _p(ks, _i32p),
                                   int(is_restricted), C.byref(v))
        return int(v.value)

    def task_info(self, sp: Space, is_restricted: bool, task):
        ks = np.ascontiguousarray(sp.k_spin, np.int32)
        kr = np.ascontiguousarray(sp.k_range, np.int64)
        s1 = np.zeros(9, np.int32)
        d1 = np.zeros(9 * sp.noab, np.int32)
        d2 = np.zeros(9 * sp.nvab, np.int32)
        cnt = np.zeros(3, np.int32)
        self.L.ref_ccsdt_task_info(sp.noa, sp.nob, sp.nva, sp.nvb, _p(kr, _i64p), _p(ks, _i32p),
                                   int(is_restricted), *[int(x) for x in task[:6]], _p(s1, _i32p),
                                   _p(d1, _i32p), _p(d2, _i32p), _p(cnt, _i32p))
        return s1, d1, d2, cnt


class ReferenceGPU(Reference):
    """The reference's own GPU task function and kernel, compiled unmodified for sm_100a against the shim
    (oracle/_ref/libccsdt_refgpu_{tc,fma}.so): the on-box GPU comparator of SURVEY.md 8(d).
    kind "tc" = K1 `fully_fused_kernel_ccsd_t_nvidia_tc_fp64` (ccsd_t_all_fused_gpu.cu:132),
    kind "fma" = K2 `revised_jk_ccsd_t_fully_fused_kernel` (ccsd_t_all_fused_nontcCuda_Hip_Sycl.cpp:95).
    Trace records: h1,h2,h3,p4,p5,p6 tile extents, thread blocks, kernel ms, unscaled task E[T], E(T)."""

    def __init__(self, kind: str = "tc"):
        super().__init__(REF_GPU_SO[kind])
        self.kind = kind
        assert int(self.L.ref_ccsdt_gpu_kernel_kind()) == (1 if kind == "tc" else 0)

    def release(self):
        self.L.ref_ccsdt_gpu_release()


# ---------------------------------------------------------------------------------------------
# Independent closed form (no tiles, no slots): the 27 equations of
# exachem/cc/ccsd_t/ccsd_t_all_fused_cpu.hpp:172-270,330-428,472-578 as whole-tensor einsums and the
# energy of :586-623, summed with the reference's restricted / unrestricted weights
# (ccsd_t_fused_driver.hpp:383-395: each unordered tile triple once, factor 2 for restricted with
#  only the aaa and aab spin cases).  Element-wise this is
#    E = sum_{i<=j<=k, a<=b<=c by TILE} factor * X  ==  sum over spin cases of sum_{all} X / (sym)
# which for data with tile-independent symmetry equals (restricted):
#    2 * [ sum_{aaa} X / 36 + sum_{aab-ordered} X / 4 ].
def closed_form_energy(noa, nob, nva, nvb, tensors, is_restricted=True):
    t1, t2 = tensors["t1"], tensors["t2"]
    Vs, Vo, Vv, evl = tensors["v2ijab"], tensors["v2ijka"], tensors["v2iabc"], tensors["evl"]
    Ot, Vt = noa + nob, nva + nvb
    eo, ev = evl[:Ot], evl[Ot:]
    osl = {1: slice(0, noa), 2: slice(noa, Ot)}
    vsl = {1: slice(0, nva), 2: slice(nva, Vt)}

    def case(si, sj, sk, sa, sb, sc):
        I, J, K, A, B, Cc = osl[si], osl[sj], osl[sk], vsl[sa], vsl[sb], vsl[sc]
        O = slice(0, Ot)
        Vf = slice(0, Vt)
        e = np.einsum
        d = np.zeros((Vt, Vt, Vt, Ot, Ot, Ot))[A, B, Cc, I, J, K]
        # d1 (sum over l): t3[a,b,c,i,j,k]
        d = d - e("abil,jklc->abcijk", t2[A, B, I, O], Vo[J, K, O, Cc])
        d = d + e("abjl,iklc->abcijk", t2[A, B, J, O], Vo[I, K, O, Cc])
        d = d - e("abkl,ijlc->abcijk", t2[A, B, K, O], Vo[I, J, O, Cc])
        d = d - e("bcil,jkla->abcijk", t2[B, Cc, I, O], Vo[J, K, O, A])
        d = d + e("bcjl,ikla->abcijk", t2[B, Cc, J, O], Vo[I, K, O, A])
        d = d - e("bckl,ijla->abcijk", t2[B, Cc, K, O], Vo[I, J, O, A])
        d = d + e("acil,jklb->abcijk", t2[A, Cc, I, O], Vo[J, K, O, B])
        d = d - e("acjl,iklb->abcijk", t2[A, Cc, J, O], Vo[I, K, O, B])
        d = d + e("ackl,ijlb->abcijk", t2[A, Cc, K, O], Vo[I, J, O, B])
        # d2 (sum over d)
        d = d - e("adij,kdbc->abcijk", t2[A, Vf, I, J], Vv[K, Vf, B, Cc])
        d = d - e("adjk,idbc->abcijk", t2[A, Vf, J, K], Vv[I, Vf, B, Cc])
        d = d + e("adik,jdbc->abcijk", t2[A, Vf, I, K], Vv[J, Vf, B, Cc])
        d = d + e("bdij,kdac->abcijk", t2[B, Vf, I, J], Vv[K, Vf, A, Cc])
        d = d + e("bdjk,idac->abcijk", t2[B, Vf, J, K], Vv[I, Vf, A, Cc])
        d = d - e("bdik,jdac->abcijk", t2[B, Vf, I, K], Vv[J, Vf, A, Cc])
        d = d - e("cdij,kdab->abcijk", t2[Cc, Vf, I, J], Vv[K, Vf, A, B])
        d = d - e("cdjk,idab->abcijk", t2[Cc, Vf, J, K], Vv[I, Vf, A, B])
        d = d + e("cdik,jdab->abcijk", t2[Cc, Vf, I, K], Vv[J, Vf, A, B])
        # s1
        s = e("ai,kjcb->abcijk", t1[A, I], Vs[K, J, Cc, B])
        s = s - e("aj,kicb->abcijk", t1[A, J], Vs[K, I, Cc, B])
        s = s + e("ak,jicb->abcijk", t1[A, K], Vs[J, I, Cc, B])
        s = s - e("bi,kjca->abcijk", t1[B, I], Vs[K, J, Cc, A])
        s = s + e("bj,kica->abcijk", t1[B, J], Vs[K, I, Cc, A])
        s = s - e("bk,jica->abcijk", t1[B, K], Vs[J, I, Cc, A])
        s = s + e("ci,kjba->abcijk", t1[Cc, I], Vs[K, J, B, A])
        s = s - e("cj,kiba->abcijk", t1[Cc, J], Vs[K, I, B, A])
        s = s + e("ck,jiba->abcijk", t1[Cc, K], Vs[J, I, B, A])
        D = (eo[I][None, None, None, :, None, None] + eo[J][None, None, None, None, :, None] +
             eo[K][None, None, None, None, None, :] - ev[A][:, None, None, None, None, None] -
             ev[B][None, :, None, None, None, None] - ev[Cc][None, None, :, None, None, None])
        return np.sum(d * d / D), np.sum(d * (d + s) / D)

    aaa = case(1, 1, 1, 1, 1, 1)
    aab = case(1, 1, 2, 1, 1, 2)
    if is_restricted:
        return (2 * (aaa[0] / 36 + aab[0] / 4), 2 * (aaa[1] / 36 + aab[1] / 4))
    abb = case(1, 2, 2, 1, 2, 2)
    bbb = case(2, 2, 2, 2, 2, 2)
    return (aaa[0] / 36 + aab[0] / 4 + abb[0] / 4 + bbb[0] / 36,
            aaa[1] / 36 + aab[1] / 4 + abb[1] / 4 + bbb[1] / 36)
