"""Host-only tests of the execution tiling (ccsdt_options.exec_tilesize): the re-cut of the caller's tiles into
box-aligned execution tiles, and the cut of an execution-tile block request into canonical storage sub-blocks.
No GPU: everything here goes through the ctx-free helpers of the C ABI."""
import itertools

import numpy as np

from exachem_b200 import _lib, driver as drv, synthetic as syn

KINDS = {0: "vo", 1: "vvoo", 2: "oovv", 3: "ooov", 4: "ovvv"}


def exec_tiles(sp, target):
    L = _lib.load()
    cap = 4096
    kr, ks, cnt = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(4, np.int32)
    n = L.ccsdt_make_exec_tiles(sp.noa, sp.nob, sp.nva, sp.nvb, np.ascontiguousarray(sp.k_range, np.int64).ctypes.data_as(_lib._i64p),
                                np.ascontiguousarray(sp.k_spin, np.int32).ctypes.data_as(_lib._i32p), target,
                                kr.ctypes.data_as(_lib._i64p), ks.ctypes.data_as(_lib._i32p), cnt.ctypes.data_as(_lib._i32p), cap)
    assert n > 0
    return drv.TiledSpace(int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3]), kr[:n].copy(), ks[:n].copy())


def split(sp, ex, tensor, bid):
    L = _lib.load()
    cap = 4096
    out = np.zeros((cap, 22), np.int64)
    b = np.zeros(4, np.uint32)
    b[:len(bid)] = bid
    cnt = np.array([ex.noa, ex.nob, ex.nva, ex.nvb], np.int32)
    n = L.ccsdt_split_request(sp.noa, sp.nob, sp.nva, sp.nvb, np.ascontiguousarray(sp.k_range, np.int64).ctypes.data_as(_lib._i64p),
                              np.ascontiguousarray(sp.k_spin, np.int32).ctypes.data_as(_lib._i32p),
                              np.ascontiguousarray(ex.k_range, np.int64).ctypes.data_as(_lib._i64p), cnt.ctypes.data_as(_lib._i32p),
                              tensor, b.ctypes.data_as(_lib._u32p), out.ctypes.data_as(_lib._i64p), cap)
    assert 0 < n <= cap
    return out[:n]


def test_exec_tiles_cover_the_same_orbitals_box_aligned():
    for (oa, ob, va, vb, ts), target in itertools.product(
            [(51, 51, 195, 195, 28), (5, 5, 19, 19, 28), (21, 21, 93, 93, 40), (60, 60, 500, 500, 16), (6, 4, 49, 51, 28),
             (3, 2, 7, 5, 2)], [-1, 8, 16, 32, 40, 48, 64]):
        sp = drv.setup_mo_space(oa, ob, va, vb, ts)
        ex = exec_tiles(sp, target)
        # same orbitals per spin block
        for g, (lo_s, hi_s, lo_e, hi_e) in enumerate([(0, sp.noa, 0, ex.noa), (sp.noa, sp.noab, ex.noa, ex.noab),
                                                       (sp.noab, sp.noab + sp.nva, ex.noab, ex.noab + ex.nva),
                                                       (sp.noab + sp.nva, sp.noab + sp.nvab, ex.noab + ex.nva, ex.noab + ex.nvab)]):
            assert sp.k_range[lo_s:hi_s].sum() == ex.k_range[lo_e:hi_e].sum() == (oa, ob, va, vb)[g]
            assert (ex.k_spin[lo_e:hi_e] == g % 2 + 1).all()
            if target > 0 or not np.array_equal(sp.k_range, ex.k_range):
                tiles = ex.k_range[lo_e:hi_e]
                gran = 2 if g < 2 else 8
                assert (tiles[:-1] % gran == 0).all()           # only the last tile of a spin block may be ragged
                assert tiles.max() <= max(target if target > 0 else 40, gran) + gran
                # padding of the whole spin block up to the box: less than one box
                assert (-(-tiles // gran) * gran).sum() - tiles.sum() < gran


def test_auto_keeps_a_tiling_the_kernel_likes_and_recuts_the_others():
    benz = drv.setup_mo_space(21, 21, 93, 93, 40)
    assert np.array_equal(exec_tiles(benz, -1).k_range, benz.k_range)          # 40,40,13: multiples of 8 but the last
    s32 = drv.setup_mo_space(60, 60, 500, 500, 32)
    assert np.array_equal(exec_tiles(s32, -1).k_range, s32.k_range)
    caf = drv.setup_mo_space(51, 51, 195, 195, 28)
    ex = exec_tiles(caf, -1)
    assert list(ex.k_range[ex.noab:ex.noab + ex.nva]) == [40, 40, 40, 40, 35]   # was 28 x 6 + 27: x1.49 padded
    s16 = drv.setup_mo_space(60, 60, 500, 500, 16)
    ex = exec_tiles(s16, -1)
    assert ex.nva == 13 and ex.k_range[ex.noab:].max() == 40                    # 32 tiles of 16 -> 13 of 40 / 32
    assert np.array_equal(exec_tiles(caf, 0).k_range, caf.k_range)             # 0 = the caller's tiles


def test_split_request_tiles_the_execution_block_exactly_with_canonical_storage_blocks():
    """every element of an execution-tile block is covered exactly once, the storage block ids are the canonically
    ordered ones the reference requests, and value(request) == sign * value(canonical block) on antisymmetric tensors"""
    orb = (5, 4, 9, 7)
    sp = drv.setup_mo_space(*orb, 3)
    T = syn.dense_all(syn.Orbitals(*orb), 5)
    dense = {0: T["t1"], 1: T["t2"], 2: T["v2ijab"], 3: T["v2ijka"], 4: T["v2iabc"]}
    so = np.concatenate([[0], np.cumsum(sp.k_range)])
    for target in (4, 8, 16):
        ex = exec_tiles(sp, target)
        eo = np.concatenate([[0], np.cumsum(ex.k_range)])
        rng = np.random.default_rng(target)
        for tensor in range(5):
            kinds = KINDS[tensor]
            for _ in range(40):
                bid = [int(rng.integers(0, ex.noab if k == "o" else ex.nvab)) for k in kinds]
                ext = [int(ex.k_range[b if k == "o" else b + ex.noab]) for b, k in zip(bid, kinds)]
                lo = [int(eo[b if k == "o" else b + ex.noab]) - (0 if k == "o" else int(eo[ex.noab])) for b, k in zip(bid, kinds)]
                want = dense[tensor][tuple(slice(l, l + e) for l, e in zip(lo, ext))]
                got = np.full(ext, np.nan)
                for q in split(sp, ex, tensor, bid):
                    cb, sign, perm = q[0:4], q[4], q[5:9]
                    s_off, e_off, ln, elems = q[9:13], q[13:17], q[17:21], q[21]
                    nd = len(kinds)
                    # canonical order of the storage block
                    if tensor == 1: assert cb[0] <= cb[1] and cb[2] <= cb[3]
                    if tensor == 2: assert cb[0] >= cb[1] and cb[2] >= cb[3]
                    if tensor == 3: assert cb[0] <= cb[1]
                    if tensor == 4: assert cb[2] <= cb[3]
                    # the canonical storage block, as Tensor::get would deliver it
                    st = [int(cb[d]) if kinds[d] == "o" else int(cb[d]) + sp.noab for d in range(nd)]
                    blo = [int(so[t]) - (0 if kinds[d] == "o" else int(so[sp.noab])) for d, t in enumerate(st)]
                    bext = [int(sp.k_range[t]) for t in st]
                    assert elems == int(np.prod(bext))
                    blk = dense[tensor][tuple(slice(l, l + e) for l, e in zip(blo, bext))]
                    # requested dim d is dim perm[d] of the canonical block
                    view = np.transpose(blk, [int(perm[d]) for d in range(nd)])
                    sub = view[tuple(slice(int(s_off[d]), int(s_off[d] + ln[d])) for d in range(nd))]
                    dst = tuple(slice(int(e_off[d]), int(e_off[d] + ln[d])) for d in range(nd))
                    assert np.isnan(got[dst]).all()                 # no element covered twice
                    got[dst] = sign * sub
                assert np.array_equal(got, want)                    # every element covered, values right
