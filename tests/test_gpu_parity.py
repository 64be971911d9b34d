"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle
(oracle/ccsdt_oracle.c, itself pinned bit-exactly to the reference's CPU path) and the committed
reference fixtures (tests/golden/ref_small.json).

Tolerance (BASELINE.json north_star): |E_gpu - E_ref| <= 1e-9 Eh absolute; FP64 throughout.  The
GPU sums in a different order (DMMA k-groups of 4, per-box partials in box order, tasks in task
order), so bit-equality with the CPU is not expected; run-to-run the GPU result IS bit-identical.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from exachem_b200 import _lib, driver as drv, synthetic as syn

pytestmark = pytest.mark.gpu
ATOL = 1e-9
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_small.json")))


def _dp(a):
    return a.ctypes.data_as(_lib._dp)


def _close(a, b):
    return abs(a - b) <= ATOL and abs(a - b) <= 1e-11 * max(1.0, abs(b))


def run_gpu(sp, T, restricted, **opts):
    ctx = drv.Context(0)
    try:
        ctx.set_options(**opts)
        ctx.set_space(sp, T["evl"], restricted)
        for tid, k in ((drv.T1, "t1"), (drv.T2, "t2"), (drv.V_IJAB, "v2ijab"), (drv.V_IJKA, "v2ijka"),
                       (drv.V_IABC, "v2iabc")):
            ctx.put_dense(tid, T[k])
        n = len(drv.enumerate_tasks(sp, restricted)[0])
        e1, e2, stats, pt = ctx.run(per_task_n=n)
        return e1, e2, stats, pt
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------
# unit probes of the hardware assumptions the kernel is built on
def test_dmma_fragment_layout():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((8, 4)), rng.standard_normal((4, 8))
    c = np.zeros((8, 8))
    assert _lib.load().ccsdt_probe_dmma_layout(0, _dp(c), _dp(a), _dp(b)) == 0
    assert np.allclose(c, a @ b, rtol=0, atol=1e-14)


def test_tma_128B_swizzle_layout():
    rows = 64
    dump = np.zeros(rows * 16)
    assert _lib.load().ccsdt_probe_tma_swizzle(0, _dp(dump), rows) == 0
    exp = np.zeros(rows * 16)
    for r in range(rows):
        for k in range(16):
            chunk = (k >> 1) ^ (r & 7)                     # 16-byte chunk XOR (row mod 8)
            exp[r * 16 + chunk * 2 + (k & 1)] = r * 16 + k
    assert np.array_equal(dump, exp)


def test_device_synthetic_generator_matches_numpy():
    orb = syn.Orbitals(3, 2, 4, 5)
    L = _lib.load()
    for tensor, dims in ((syn.T1, (orb.Vt, orb.Ot, 1, 1)), (syn.T2, (orb.Vt, orb.Vt, orb.Ot, orb.Ot)),
                         (syn.V_IJAB, (orb.Ot, orb.Ot, orb.Vt, orb.Vt)), (syn.V_IJKA, (orb.Ot, orb.Ot, orb.Ot, orb.Vt)),
                         (syn.V_IABC, (orb.Ot, orb.Vt, orb.Vt, orb.Vt))):
        lo = np.zeros(4, np.int64)
        n = np.array(dims, np.int64)
        out = np.zeros(int(np.prod(n)))
        assert L.ccsdt_synth_block(0, 77, tensor, 3, 2, 4, 5, lo.ctypes.data_as(_lib._i64p),
                                   n.ctypes.data_as(_lib._i64p), _dp(out)) == 0
        ref = syn.dense(orb, 77, tensor).ravel()
        assert np.array_equal(out, ref)                    # bit-identical


def test_fp64_peak_probe_runs():
    L = _lib.load()
    tf, ms = C.c_double(0), C.c_double(0)
    for use_dmma in (0, 1):
        assert L.ccsdt_probe_fp64_peak(0, use_dmma, 2000, C.byref(tf), C.byref(ms)) == 0
        assert tf.value > 1.0
        print(f"fp64 peak probe dmma={use_dmma}: {tf.value:.2f} TFLOP/s ({ms.value:.3f} ms)")


# ------------------------------------------------------------------------------------------------
# parity against the committed reference fixtures and the oracle
BOXES = [(1, 1, 1), (1, 1, 2), (2, 1, 1), (1, 1, 3), (1, 3, 1)]


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("kernel", ["simple", "dmma"])
def test_energy_matches_reference_fixture(name, kernel):
    g = GOLD[name]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    e1, e2, stats, pt = run_gpu(sp, T, g["restricted"],
                                kernel={"simple": drv.KERNEL_SIMPLE, "dmma": drv.KERNEL_DMMA}[kernel])
    assert _close(e1, float(g["energy1"])) and _close(e2, float(g["energy2"]))
    run1 = np.cumsum(pt[:, 0])
    assert np.allclose(run1, [float(x) for x in g["running_e1"]], rtol=0, atol=ATOL)   # task by task
    assert stats["tasks_run"] == len(g["tasks"])
    assert stats["counted_flops"] == g["total_num_ops"]


@pytest.mark.parametrize("sub", BOXES)
@pytest.mark.parametrize("cfg", [(4, 4, 6, 6, 3, True, 11), (5, 5, 11, 11, 8, True, 99), (3, 3, 5, 5, 2, False, 8),
                                 (6, 6, 17, 17, 9, True, 21), (9, 9, 10, 10, 10, True, 5)])
def test_dmma_kernel_all_box_shapes(orc, cfg, sub):
    oa, ob, va, vb, ts, restricted, seed = cfg
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
    ref = orc.run(osp, T, restricted, per_task=True)
    e1, e2, _, pt = run_gpu(sp, T, restricted, sub=sub)
    assert _close(e1, ref[0]) and _close(e2, ref[1])
    assert np.allclose(pt, ref[2], rtol=0, atol=ATOL)


@pytest.mark.parametrize("opts", [dict(stages=2), dict(sub=(1, 1, 1), ctas_per_sm=2), dict(sub=(1, 1, 2), stages=3),
                                  dict(sub=(1, 1, 1), ctas_per_sm=1, stages=2)])
def test_ring_depth_and_occupancy_variants(orc, opts):
    """the TMA ring must be correct at its minimum depth and at every CTA-per-SM setting"""
    oa, ob, va, vb, ts = 6, 6, 17, 17, 9
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 21)
    ref = orc.run(osp, T, True)
    e1, e2, _, _ = run_gpu(sp, T, True, **opts)
    assert _close(e1, ref[0]) and _close(e2, ref[1])


def test_ragged_tiles_and_single_orbital_tiles(orc):
    """tile extents 1..7 (every partial-box case), incl. a 1-orbital occupied tile"""
    oa = ob = 5
    va = vb = 13
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, 4), orc.tiles(oa, ob, va, vb, 4)   # occ [4,1], virt [4,4,4,1]
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 3)
    ref = orc.run(osp, T, True)
    e1, e2, _, _ = run_gpu(sp, T, True)
    assert _close(e1, ref[0]) and _close(e2, ref[1])


def test_unequal_alpha_beta_counts(orc):
    sp, osp = drv.setup_mo_space(4, 3, 6, 7, 4), orc.tiles(4, 3, 6, 7, 4)
    T = syn.dense_all(syn.Orbitals(4, 3, 6, 7), 17)
    ref = orc.run(osp, T, False)
    e1, e2, _, _ = run_gpu(sp, T, False)
    assert _close(e1, ref[0]) and _close(e2, ref[1])


def test_bitwise_reproducible_and_overlap_independent():
    sp = drv.setup_mo_space(6, 6, 17, 17, 9)
    T = syn.dense_all(syn.Orbitals(6, 6, 17, 17), 21)
    a = run_gpu(sp, T, True)
    b = run_gpu(sp, T, True)
    c = run_gpu(sp, T, True, overlap=0)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[3], b[3])
    assert a[0] == c[0] and a[1] == c[1]


def test_block_upload_and_fetch_callback_paths(orc):
    oa, ob, va, vb, ts = 4, 4, 6, 6, 3
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 1234)
    ref = orc.run(osp, T, True)
    off = sp.k_offset
    names = {drv.T1: ("t1", "vo"), drv.T2: ("t2", "vvoo"), drv.V_IJAB: ("v2ijab", "oovv"),
             drv.V_IJKA: ("v2ijka", "ooov"), drv.V_IABC: ("v2iabc", "ovvv")}
    calls = []

    def get_block(tensor, bid):
        name, kinds = names[tensor]
        sl = []
        for k, b in zip(kinds, bid):
            t = b if k == "o" else b + sp.noab
            base = 0 if k == "o" else off[sp.noab]
            sl.append(slice(off[t] - base, off[t + 1] - base))
        calls.append((tensor, bid))
        return np.ascontiguousarray(T[name][tuple(sl)])

    class Blocks:  # what a TAMM Tensor looks like to the driver: .get(block id)
        def __init__(self, tensor):
            self.tensor = tensor

        def get(self, bid):
            return get_block(self.tensor, bid)

    d = drv.CCSD_T_Fused_Driver(device=0)
    e1, e2, _, _ = d.execute(None, None, sp.k_spin, sp, Blocks(drv.T1), Blocks(drv.T2),
                             {"v2ijab": Blocks(drv.V_IJAB), "v2ijka": Blocks(drv.V_IJKA),
                              "v2iabc": Blocks(drv.V_IABC)}, T["evl"], 0.0, True)
    assert _close(e1, ref[0]) and _close(e2, ref[1])
    assert d.last_stats["blocks_fetched"] == len(calls) > 0
    assert len(set(calls)) == len(calls)      # every block crosses PCIe once: the HBM store caches it
    # the reference only ever requests canonically ordered T2 blocks (p_lo<=p_hi, h_lo<=h_hi)
    assert all(b[0] <= b[1] and b[2] <= b[3] for t, b in calls if t == drv.T2)


def test_device_generated_tensors_equal_uploaded_ones():
    oa, ob, va, vb, ts, seed = 5, 5, 11, 11, 8, 99
    sp = drv.setup_mo_space(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
    a = run_gpu(sp, T, True)
    ctx = drv.Context(0)
    try:
        ctx.set_space(sp, T["evl"], True)
        ctx.set_synthetic(seed)
        e1, e2, _, _ = ctx.run()
    finally:
        ctx.close()
    assert e1 == a[0] and e2 == a[1]


def test_rank_partials_sum_to_total():
    sp = drv.setup_mo_space(6, 6, 17, 17, 5)
    T = syn.dense_all(syn.Orbitals(6, 6, 17, 17), 2)
    tot = run_gpu(sp, T, True)
    parts = [run_gpu(sp, T, True, rank=r, nranks=3) for r in range(3)]
    assert abs(sum(p[0] for p in parts) - tot[0]) < 1e-12 and abs(sum(p[1] for p in parts) - tot[1]) < 1e-12
    assert sum(p[2]["tasks_run"] for p in parts) == tot[2]["tasks_run"]


def test_tiling_invariance_medium(orc):
    """size-independent property at a size the CPU oracle would need minutes for: the energy does
    not depend on the tile size (only possible if layouts, signs, factors and masks are all right)"""
    oa = ob = 10
    va = vb = 24
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 7)
    es = [run_gpu(drv.setup_mo_space(oa, ob, va, vb, ts), T, True)[:2] for ts in (24, 12, 7)]
    for e in es[1:]:
        assert abs(e[0] - es[0][0]) < 1e-9 * max(1, abs(es[0][0])) and abs(e[1] - es[0][1]) < 1e-9 * max(1, abs(es[0][1]))


def test_dynamic_task_counter_two_claimants():
    """two contexts claim tasks from one shared counter concurrently (the AtomicCounterGA role): every task
    runs exactly once and the partials sum to the total"""
    import threading
    sp = drv.setup_mo_space(6, 6, 17, 17, 5)
    T = syn.dense_all(syn.Orbitals(6, 6, 17, 17), 2)
    tot = run_gpu(sp, T, True)
    n = len(drv.enumerate_tasks(sp, True)[0])
    counter = C.c_int64(0)
    out = [None, None]

    def worker(i):
        ctx = drv.Context(0)
        try:
            ctx.set_space(sp, T["evl"], True)
            for tid, k in ((drv.T1, "t1"), (drv.T2, "t2"), (drv.V_IJAB, "v2ijab"), (drv.V_IJKA, "v2ijka"),
                           (drv.V_IABC, "v2iabc")):
                ctx.put_dense(tid, T[k])
            ctx.set_task_counter(C.addressof(counter))
            out[i] = ctx.run(per_task_n=n)
        finally:
            ctx.close()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert out[0] is not None and out[1] is not None
    assert out[0][2]["tasks_run"] + out[1][2]["tasks_run"] == n and counter.value >= n
    pt = out[0][3] + out[1][3]
    assert np.array_equal(pt, tot[3])                      # each task computed once, bit-identical to the static run
    assert abs(out[0][0] + out[1][0] - tot[0]) < 1e-12 and abs(out[0][1] + out[1][1] - tot[1]) < 1e-12


def test_full_width_tiles_vs_oracle(orc):
    """ccsdt_tilesize 40 (ExaChem's default): five 8-wide particle boxes per index, whole job against the oracle"""
    oa = ob = 6
    va = vb = 40
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, 40), orc.tiles(oa, ob, va, vb, 40)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 40)
    ref = orc.run(osp, T, True, per_task=True)
    e1, e2, st, pt = run_gpu(sp, T, True)
    assert _close(e1, ref[0]) and _close(e2, ref[1])
    assert np.allclose(pt, ref[2], rtol=1e-11, atol=ATOL)
    assert st["counted_flops"] == orc.count_ops(osp, True)


def test_headline_size_task_is_box_shape_independent():
    """BASELINE full size, (nocc, nvir) = (60, 500) at ccsdt_tilesize 32 on device-generated tensors: one
    32^6 task (1.08e13 counted flops, 262 144 CTA boxes) must give the same energies for every CTA box
    shape / ring depth (different boxes, bricks and partial sums; same elements), and a ragged task must too."""
    no, nv, ts = 60, 500, 32
    sp = drv.setup_mo_space(no, no, nv, nv, ts)
    evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
    tasks, _, _ = drv.enumerate_tasks(sp, True)
    ragged = next(i for i, t in enumerate(tasks) if sp.k_range[t[5]] == 20 and sp.k_range[t[0]] == 32 and sp.k_range[t[2]] == 28)
    res = {}
    for name, opts in (("default", {}), ("box_1_1_2", dict(sub=(1, 1, 2))), ("two_stage", dict(stages=2, ctas_per_sm=2))):
        ctx = drv.Context(0)
        try:
            ctx.set_options(**opts)
            ctx.set_space(sp, evl, True)
            ctx.set_synthetic(1234)
            res[name] = ctx.run_tasks([0, ragged], per_task=True)
        finally:
            ctx.close()
    ref = res["default"]
    assert ref[2]["tasks_run"] == 2 and ref[2]["counted_flops"] > 1.0e13
    assert np.all(np.isfinite(ref[3])) and np.all(ref[3] < 0)          # D < 0 everywhere: E[T] contributions are negative
    for name in ("box_1_1_2", "two_stage"):
        assert np.allclose(res[name][3], ref[3], rtol=1e-11, atol=0), name


def test_task_list_subset_and_order_independent():
    """ccsdt_run_tasks: per-task energies do not depend on which other tasks run or in what order"""
    sp = drv.setup_mo_space(6, 6, 17, 17, 5)
    T = syn.dense_all(syn.Orbitals(6, 6, 17, 17), 2)
    full = run_gpu(sp, T, True)
    n = len(drv.enumerate_tasks(sp, True)[0])
    ids = np.random.default_rng(0).permutation(n)[: n // 3]
    ctx = drv.Context(0)
    try:
        ctx.set_space(sp, T["evl"], True)
        for tid, k in ((drv.T1, "t1"), (drv.T2, "t2"), (drv.V_IJAB, "v2ijab"), (drv.V_IJKA, "v2ijka"), (drv.V_IABC, "v2iabc")):
            ctx.put_dense(tid, T[k])
        e1, e2, st, pt = ctx.run_tasks(ids, per_task=True)
    finally:
        ctx.close()
    assert np.array_equal(pt, full[3][ids]) and st["tasks_run"] == len(ids)
    assert abs(e1 - full[3][ids, 0].sum()) < 1e-12


@pytest.mark.parametrize("cfg", [(6, 6, 17, 17, 40, True, 3),      # one tile per spin block: every ααα task is 6x6-fold symmetric
                                 (9, 9, 21, 21, 8, True, 4),       # mixed: equal pairs, equal triples, distinct tiles, ragged tails
                                 (5, 3, 9, 12, 4, False, 5),       # unrestricted, unequal alpha/beta counts (bbb / abb tasks too)
                                 (7, 7, 13, 13, 5, True, 6)])
def test_symmetry_reduction_equals_full_evaluation(orc, cfg):
    """options.symmetry (default on) skips the CTA boxes that are mirror images under exchange of indices whose
    TILES coincide and weights the rest; per task it must give what the full evaluation and the oracle give."""
    oa, ob, va, vb, ts, restricted, seed = cfg
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
    ref = orc.run(osp, T, restricted, per_task=True)
    on = run_gpu(sp, T, restricted, symmetry=1)
    off = run_gpu(sp, T, restricted, symmetry=0)
    for r in (on, off):
        assert _close(r[0], ref[0]) and _close(r[1], ref[1])
        assert np.allclose(r[3], ref[2], rtol=1e-11, atol=ATOL)
    assert off[2]["evaluated_flops"] == off[2]["counted_flops"] == on[2]["counted_flops"]
    assert on[2]["evaluated_flops"] < 0.8 * on[2]["counted_flops"]
    # other CTA boxes (hole box 2,2,4: the h2~h3 symmetry is then not usable) and the diagnostic kernel
    alt = run_gpu(sp, T, restricted, symmetry=1, sub=(1, 1, 2))
    assert np.allclose(alt[3], ref[2], rtol=1e-11, atol=ATOL)


def test_symmetry_reduction_headline_shape():
    """(60,500) ts32 on device-generated tensors: task 0 (tiles 0,0,0 | 4,4,4: 36-fold symmetric) and an aab task
    with h1b == h2b, symmetry on vs off"""
    no, nv, ts = 60, 500, 32
    sp = drv.setup_mo_space(no, no, nv, nv, ts)
    evl = syn.Orbitals(no, no, nv, nv).orbital_energies()
    tasks, _, _ = drv.enumerate_tasks(sp, True)
    aab = next(i for i, t in enumerate(tasks) if t[0] == t[1] and sp.k_spin[t[2]] == 2 and t[3] != t[4])
    res = {}
    for s in (1, 0):
        ctx = drv.Context(0)
        try:
            ctx.set_options(symmetry=s)
            ctx.set_space(sp, evl, True)
            ctx.set_synthetic(1234)
            res[s] = ctx.run_tasks([0, aab], per_task=True)
        finally:
            ctx.close()
    assert np.allclose(res[1][3], res[0][3], rtol=1e-11, atol=0)
    assert res[1][2]["evaluated_flops"] < 0.35 * res[0][2]["evaluated_flops"]
    assert res[1][2]["seconds_kernel"] < 0.5 * res[0][2]["seconds_kernel"]


def _spin_orbital_cholesky(oa, ob, va, vb, ncv, seed):
    """random Cholesky-like vectors L[p,q,c] = L[q,p,c] over spin orbitals | occ a | occ b | virt a | virt b |, zero
    between different spins (what ExaChem's cholVpr looks like), and the V2 tensors setupV2Tensors derives from them"""
    rng = np.random.default_rng(seed)
    spin = np.array([0] * oa + [1] * ob + [0] * va + [1] * vb)
    n = len(spin)
    L = rng.uniform(-1, 1, (n, n, ncv)) * 0.3
    L = 0.5 * (L + L.transpose(1, 0, 2)) * (spin[:, None] == spin[None, :])[:, :, None]
    O = oa + ob
    Loo, Lov, Lvv = L[:O, :O], L[:O, O:], L[O:, O:]
    v2ijab = np.einsum("iac,jbc->ijab", Lov, Lov) - np.einsum("ibc,jac->ijab", Lov, Lov)     # v2tensors.cpp:68-69
    v2ijka = np.einsum("ikc,jac->ijka", Loo, Lov) - np.einsum("jkc,iac->ijka", Loo, Lov)     # v2tensors.cpp:77-78
    v2iabc = np.einsum("ibc,adc->iabd", Lov, Lvv) - np.einsum("idc,abc->iabd", Lov, Lvv)     # v2tensors.cpp:85-86
    return L, v2ijab, v2ijka, v2iabc


@pytest.mark.parametrize("cfg", [(4, 4, 7, 7, 3, 13, 1), (5, 3, 9, 6, 4, 20, 2)])
def test_v2_tensors_from_cholesky_vectors_on_the_device(orc, cfg):
    """ccsdt_put_cholesky (row f2: setupV2Tensors on the GPU) must give the energies of the same V2 tensors built on
    the host with the reference's formulas and uploaded with ccsdt_put_dense, and the oracle's"""
    oa, ob, va, vb, ts, ncv, seed = cfg
    restricted = oa == ob
    sp, osp = drv.setup_mo_space(oa, ob, va, vb, ts), orc.tiles(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
    L, T["v2ijab"], T["v2ijka"], T["v2iabc"] = _spin_orbital_cholesky(oa, ob, va, vb, ncv, seed)
    ref = orc.run(osp, T, restricted, per_task=True)
    dense = run_gpu(sp, T, restricted)
    ctx = drv.Context(0)
    try:
        ctx.set_space(sp, T["evl"], restricted)
        ctx.put_dense(drv.T1, T["t1"])
        ctx.put_dense(drv.T2, T["t2"])
        ctx.put_cholesky(L)
        n = len(drv.enumerate_tasks(sp, restricted)[0])
        e1, e2, st, pt = ctx.run(per_task_n=n)
    finally:
        ctx.close()
    assert _close(e1, ref[0]) and _close(e2, ref[1])
    assert np.allclose(pt, ref[2], rtol=1e-11, atol=ATOL)
    assert np.allclose(pt, dense[3], rtol=1e-12, atol=1e-12)


def test_random_shapes_symmetric_dmma_kernel_equals_elementwise_kernel():
    """30 seeded random problems (odd and tiny tiles, one-orbital tiles, unequal alpha/beta, restricted and not): the
    product kernel with the symmetry reduction against the diagnostic one-thread-per-element kernel, task by task"""
    rng = np.random.default_rng(20261017)
    for trial in range(30):
        restricted = bool(rng.integers(0, 2))
        oa = int(rng.integers(1, 8))
        ob = oa if restricted else int(rng.integers(1, 8))
        va = int(rng.integers(2, 20))
        vb = va if restricted else int(rng.integers(2, 20))
        ts = int(rng.integers(1, 12))
        sp = drv.setup_mo_space(oa, ob, va, vb, ts)
        if len(drv.enumerate_tasks(sp, restricted)[0]) == 0:
            continue
        T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 1000 + trial)
        a = run_gpu(sp, T, restricted, symmetry=1)
        b = run_gpu(sp, T, restricted, kernel=drv.KERNEL_SIMPLE)
        assert np.allclose(a[3], b[3], rtol=1e-11, atol=1e-13), (trial, oa, ob, va, vb, ts, restricted)
        assert a[2]["counted_flops"] == b[2]["counted_flops"]


@pytest.mark.parametrize("cfg", [(6, 6, 17, 17, 5, True), (5, 3, 9, 12, 4, False)])
def test_asynchronous_dense_upload_gives_identical_energies(cfg):
    """ccsdt_put_dense_async: copies in flight while the all-alpha tasks already run; same per-task energies as the
    synchronous upload, bit for bit, and the upload bytes show up in the run's stats"""
    oa, ob, va, vb, ts, restricted = cfg
    sp = drv.setup_mo_space(oa, ob, va, vb, ts)
    T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), 77)
    sync = run_gpu(sp, T, restricted)
    n = len(drv.enumerate_tasks(sp, restricted)[0])
    ctx = drv.Context(0)
    try:
        ctx.set_space(sp, T["evl"], restricted)
        for rep in range(2):          # second pass re-uploads into the same device buffers
            for tid, k in ((drv.T1, "t1"), (drv.T2, "t2"), (drv.V_IJAB, "v2ijab"), (drv.V_IJKA, "v2ijka"), (drv.V_IABC, "v2iabc")):
                ctx.put_dense(tid, T[k], async_=True)
            e1, e2, st, pt = ctx.run(per_task_n=n)
            assert np.array_equal(pt, sync[3]) and e1 == sync[0] and e2 == sync[1]
            assert 0 < st["h2d_bytes"] < sum(T[k].nbytes for k in ("t1", "t2", "v2ijab", "v2ijka", "v2iabc"))
    finally:
        ctx.close()


def test_n60v500_full_tile_task_matches_the_reference_cpu_golden():
    """Kernel task 0 of the north_star problem -- (nocc, nvir) = (60, 500), ccsdt_tilesize 32, all six tiles full: the 32^6 task,
    the largest tile the reference CPU kernel can index -- against the energy the REFERENCE's own CPU path gave on the same
    procedural tensors (tests/golden/n60v500_ts32_task0.json, made by tests/golden/make_n60v500_task0_golden.py in two hours
    of CPU).  The synthetic energies are of order 4e5, so the 1e-9 Eh bar of real molecules (|E(T)| ~ 1e-2 Eh, i.e. 1e-7
    relative) is applied as a relative bound, tightened to 1e-11; with the symmetry reduction (36-fold here) and without."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "n60v500_ts32_task0.json")))
    sp = drv.setup_mo_space(60, 60, 500, 500, 32)
    evl = syn.Orbitals(60, 60, 500, 500).orbital_energies()
    for symmetry in (1, 0):
        ctx = drv.Context(0)
        try:
            ctx.set_options(symmetry=symmetry)
            ctx.set_space(sp, evl, True)
            ctx.set_synthetic(1234)
            tasks, _, _ = drv.enumerate_tasks(sp, True)
            assert list(tasks[0][:6]) == g["task"]
            e1, e2, st, _ = ctx.run_tasks([0])
        finally:
            ctx.close()
        r1, r2 = abs(e1 - g["E[T]"]) / abs(g["E[T]"]), abs(e2 - g["E(T)"]) / abs(g["E(T)"])
        print(f"(60,500) task 0, symmetry {symmetry}: relative deviation from the reference CPU kernel {r1:.2e} / {r2:.2e}")
        assert r1 < 1e-11 and r2 < 1e-11, (e1, e2, g)

