"""GPU tests of the operand store and the execution tiling: asynchronous block fetch (pinned ring, prefetch), LRU
eviction under a forced small budget, blocks supplied by put_block, the parked context, and exec_tilesize (the kernel
works on box-aligned execution tiles while blocks are still requested in the caller's tiling).

Bars: 1e-9 Eh against the committed reference fixtures / the oracle; results of the SAME execution tiling must be
bit-identical whatever the way the operands arrived (dense, put_block, fetch, evicted and re-fetched, prefetched)."""
import json
import os

import numpy as np
import pytest

from exachem_b200 import _lib, driver as drv, synthetic as syn

pytestmark = pytest.mark.gpu
ATOL = 1e-9
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_small.json")))
NAMES = {drv.T1: ("t1", "vo"), drv.T2: ("t2", "vvoo"), drv.V_IJAB: ("v2ijab", "oovv"), drv.V_IJKA: ("v2ijka", "ooov"),
         drv.V_IABC: ("v2iabc", "ovvv")}


def _close(a, b):
    return abs(a - b) <= ATOL and abs(a - b) <= 1e-11 * max(1.0, abs(b))


def block_of(sp, T, tensor, bid):
    name, kinds = NAMES[tensor]
    off, sl = sp.k_offset, []
    for k, b in zip(kinds, bid):
        t = b if k == "o" else b + sp.noab
        base = 0 if k == "o" else off[sp.noab]
        sl.append(slice(off[t] - base, off[t + 1] - base))
    return np.ascontiguousarray(T[name][tuple(sl)])


def is_canonical(tensor, bid):
    if tensor == drv.T2:
        return bid[0] <= bid[1] and bid[2] <= bid[3]
    if tensor == drv.V_IJAB:
        return bid[0] >= bid[1] and bid[2] >= bid[3]
    if tensor == drv.V_IJKA:
        return bid[0] <= bid[1]
    if tensor == drv.V_IABC:
        return bid[2] <= bid[3]
    return True


def run(sp, T, restricted, how="dense", calls=None, **opts):
    ctx = drv.Context(0)
    try:
        ctx.set_options(**opts)
        ctx.set_space(sp, T["evl"], restricted)
        if how == "dense":
            for tid, (k, _) in NAMES.items():
                ctx.put_dense(tid, T[k])
        elif how == "fetch":
            def get(tensor, bid):
                if calls is not None:
                    calls.append((tensor, tuple(bid)))
                return block_of(sp, T, tensor, bid)
            ctx.set_fetch(get)
        elif how == "put_block":
            for tid, (_, kinds) in NAMES.items():
                rng = [range(sp.noab if k == "o" else sp.nvab) for k in kinds]
                import itertools
                for bid in itertools.product(*rng):
                    if is_canonical(tid, bid):
                        ctx.put_block(tid, bid, block_of(sp, T, tid, bid))
        e1, e2, st, _ = ctx.run()
        return e1, e2, st
    finally:
        ctx.close()


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("target", [-1, 8, 16])
def test_execution_tiling_reproduces_the_reference_energy(name, target):
    """the total does not depend on the execution tiling (dense operands)"""
    g = GOLD[name]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    e1, e2, st = run(sp, T, g["restricted"], exec_tilesize=target)
    assert _close(e1, float(g["energy1"])) and _close(e2, float(g["energy2"]))


@pytest.mark.parametrize("name", ["o4v6_ts3", "o3v9_ts4", "uhf_o3o2_v5v6_ts3", "h2o_shape_ts7"])
@pytest.mark.parametrize("target", [0, 8, 16])
def test_fetch_in_the_callers_tiling_under_an_execution_tiling(name, target):
    """blocks are requested in the CALLER's tiling, only canonically ordered ones, each once; the result is bit-identical
    to the dense run of the same execution tiling"""
    g = GOLD[name]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    calls = []
    e1, e2, st = run(sp, T, g["restricted"], how="fetch", calls=calls, exec_tilesize=target)
    d1, d2, _ = run(sp, T, g["restricted"], exec_tilesize=target)
    assert (e1, e2) == (d1, d2)
    assert _close(e1, float(g["energy1"])) and _close(e2, float(g["energy2"]))
    assert st["blocks_fetched"] == len(calls) == len(set(calls)) > 0
    assert all(is_canonical(t, b) for t, b in calls)
    p1, p2, _ = run(sp, T, g["restricted"], how="put_block", exec_tilesize=target)
    assert (p1, p2) == (d1, d2)


def test_eviction_under_a_forced_small_budget():
    """a budget far below the working set: blocks are evicted and fetched again, the energies do not change by a bit"""
    g = GOLD["h2o_shape_ts7"]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    calls, calls_small = [], []
    e1, e2, st = run(sp, T, True, how="fetch", calls=calls)
    assert st["blocks_evicted"] == 0
    s1, s2, st_small = run(sp, T, True, how="fetch", calls=calls_small, block_budget_bytes=64 * 1024)
    assert (s1, s2) == (e1, e2)
    assert st_small["blocks_evicted"] > 0 and len(calls_small) > len(calls)      # evicted blocks came back
    assert _close(s1, float(g["energy1"])) and _close(s2, float(g["energy2"]))
    # put_block'ed blocks are never evicted (no callback could bring them back)
    p1, p2, st_p = run(sp, T, True, how="put_block", block_budget_bytes=64 * 1024)
    assert (p1, p2) == (e1, e2) and st_p["blocks_evicted"] == 0


def test_prefetch_and_execution_order_do_not_change_the_result():
    g = GOLD["o3v9_ts4"]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    a = run(sp, T, True, how="fetch", prefetch_tasks=-1)
    b = run(sp, T, True, how="fetch", prefetch_tasks=8)
    c = run(sp, T, True)
    assert a[:2] == b[:2] == c[:2]
    assert a[2]["blocks_fetched"] == b[2]["blocks_fetched"]


def test_parked_context_is_reused_and_clean():
    """ccsdt_destroy parks the device resources; the next create must behave like a fresh context"""
    L = _lib.load()
    L.ccsdt_release_cached()
    ga, gb = GOLD["o5v11_ts8"], GOLD["uhf_o3o2_v5v6_ts3"]
    out = []
    for g in (ga, gb, ga):                       # different spaces back to back through the same parked context
        sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
        T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
        e1, e2, st = run(sp, T, g["restricted"], how="fetch")
        assert _close(e1, float(g["energy1"])) and _close(e2, float(g["energy2"]))
        assert st["blocks_fetched"] > 0          # nothing survived from the previous owner of the context
        out.append((e1, e2))
    assert out[0] == out[2]
    # a parked context forgets options, fetch callback and operands
    ctx = drv.Context(0)
    try:
        sp = drv.setup_mo_space(ga["noa"], ga["nob"], ga["nva"], ga["nvb"], ga["tilesize"])
        T = syn.dense_all(syn.Orbitals(ga["noa"], ga["nob"], ga["nva"], ga["nvb"]), ga["seed"])
        ctx.set_space(sp, T["evl"], True)
        with pytest.raises(drv.CcsdtError, match="never supplied"):
            ctx.run()
    finally:
        ctx.close()
    L.ccsdt_release_cached()
    e1, e2, _ = run(sp, T, True)                 # and a really fresh one after release
    assert (e1, e2) == out[0]


def test_clear_blocks_makes_the_next_run_cold():
    g = GOLD["o4v6_ts4"]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    ctx = drv.Context(0)
    try:
        ctx.set_space(sp, T["evl"], True)
        ctx.set_fetch(lambda tensor, bid: block_of(sp, T, tensor, bid))
        a = ctx.run()
        b = ctx.run()                            # warm: everything resident
        ctx.clear_blocks()
        c = ctx.run()
        assert a[:2] == b[:2] == c[:2]
        assert a[2]["blocks_fetched"] > 0 and b[2]["blocks_fetched"] == 0 and c[2]["blocks_fetched"] == a[2]["blocks_fetched"]
    finally:
        ctx.close()


def test_rank_without_tasks_returns_zero_and_waits_for_uploads():
    """nranks larger than the task count: the idle rank still synchronises the asynchronous uploads"""
    g = GOLD["o4v6_whole"]                        # 2 kernel tasks
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"])
    tot = np.zeros(2)
    for r in range(4):
        ctx = drv.Context(0)
        try:
            ctx.set_options(rank=r, nranks=4)
            ctx.set_space(sp, T["evl"], True)
            for tid, (k, _) in NAMES.items():
                ctx.put_dense(tid, T[k], async_=True)
            e1, e2, st, _ = ctx.run()
            tot += (e1, e2)
            if st["tasks_run"] == 0:
                assert (e1, e2) == (0.0, 0.0)
        finally:
            ctx.close()
    assert _close(tot[0], float(g["energy1"])) and _close(tot[1], float(g["energy2"]))


def test_executed_flops_account_for_padding():
    """executed_flops >= evaluated_flops; equal when every tile is a multiple of the box and K of 4"""
    sp = drv.setup_mo_space(4, 4, 16, 16, 8)     # holes 4, particles 8,8: no padding at all (K = 4 and 16)
    T = syn.dense_all(syn.Orbitals(4, 4, 16, 16), 3)
    _, _, st = run(sp, T, True, symmetry=0)
    s1 = st["counted_flops"] - st["executed_flops"]          # the s1 outer products are not DMMAs
    assert 0 < s1 < 0.2 * st["counted_flops"]
    sp2 = drv.setup_mo_space(5, 5, 19, 19, 28)   # ragged everything
    T2 = syn.dense_all(syn.Orbitals(5, 5, 19, 19), 3)
    _, _, st2 = run(sp2, T2, True, symmetry=0)
    assert st2["executed_flops"] > 1.3 * st2["evaluated_flops"]


def test_single_rank_nccl_allreduce_through_the_c_abi():
    import ctypes as C
    L = _lib.load()
    uid = C.create_string_buffer(128)
    if L.ccsdt_comm_unique_id(uid) != 0:
        pytest.skip("libnccl.so.2 not loadable")
    ctx = drv.Context(0)
    try:
        ctx.comm_init(uid.raw, 0, 1)
        assert ctx.comm_allreduce(1.5, -2.25) == (1.5, -2.25)
    finally:
        ctx.close()


@pytest.mark.parametrize("how", ["dense", "fetch"])
def test_symmetry_reduction_refuses_operands_that_are_not_antisymmetric(how, orc):
    """options.symmetry skips mirror-image boxes, which is only right for antisymmetric T2 / V2: the device check
    (options.check_symmetry, default on) must turn a violation into an error, and symmetry = 0 must still give the
    reference's answer on the very same tensors"""
    g = GOLD["o4v6_ts4"]
    sp = drv.setup_mo_space(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"])
    T = dict(syn.dense_all(syn.Orbitals(g["noa"], g["nob"], g["nva"], g["nvb"]), g["seed"]))
    ok = run(sp, T, True, how=how)                              # antisymmetric as generated: passes
    assert _close(ok[0], float(g["energy1"]))
    bad = dict(T)
    bad["t2"] = T["t2"].copy()
    bad["t2"][1, 0, 1, 0] += 0.25                               # T2[a,b,i,j] != -T2[b,a,i,j] inside a diagonal block
    with pytest.raises(drv.CcsdtError, match="not antisymmetric"):
        run(sp, bad, True, how=how)
    e1, e2, _ = run(sp, bad, True, how=how, symmetry=0)         # every element evaluated, as the reference does
    r1, r2 = orc.run(orc.tiles(g["noa"], g["nob"], g["nva"], g["nvb"], g["tilesize"]), bad, True)
    assert _close(e1, r1) and _close(e2, r2)
    e1, e2, _ = run(sp, bad, True, how=how, check_symmetry=-1)  # the check can be switched off (the answer is then the caller's risk)
    assert np.isfinite(e1) and np.isfinite(e2)

