"""Synthetic spin-orbital amplitudes / integrals for the (T) path.

Counter-based: every element is a pure function of (seed, tensor id, canonical index tuple), so any
block can be produced anywhere -- numpy here, C++ in oracle/ref_driver.cpp (`synth_value`), CUDA in
exachem_b200/csrc/ccsdt_kernels.cu (`synth_value`) -- without materialising the tensor.  The three
implementations are checked against each other in tests/.

Tensors follow the index order of the TAMM tensors the reference driver consumes
(exachem/cc/ccsd_t/ccsd_t.cpp:66-69): T1[V,O], T2[V,V,O,O], v2ijab[O,O,V,V], v2ijka[O,O,O,V],
v2iabc[O,V,V,V]; orbitals are ordered | occ alpha | occ beta | and | virt alpha | virt beta |,
spin alpha = 1, beta = 2 (exachem/cholesky/cholesky_2e.cpp:192-226).
Symmetries: T2 antisymmetric in (a,b) and (i,j); v2ijab in (i,j),(a,b); v2ijka in (i,j); v2iabc in
(b,c); every tensor vanishes unless the spin sums of its index pairs agree.
"""
from __future__ import annotations

import numpy as np

T1, T2, V_IJAB, V_IJKA, V_IABC = 0, 1, 2, 3, 4
SCALE = {T1: 0.05, T2: 0.1, V_IJAB: 0.1, V_IJKA: 0.1, V_IABC: 0.1}
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def raw_value(seed: int, tensor: int, i0, i1, i2, i3) -> np.ndarray:
    """uniform(-1,1) * SCALE[tensor] keyed on (seed, tensor, i0..i3); indices < 65536."""
    i0, i1, i2, i3 = (np.asarray(a, dtype=np.uint64) for a in (i0, i1, i2, i3))
    key = (i0 << np.uint64(48)) | (i1 << np.uint64(32)) | (i2 << np.uint64(16)) | i3
    with np.errstate(over="ignore"):
        s = _mix(np.uint64(seed) ^ (np.uint64(tensor + 1) * np.uint64(0xD1B54A32D192ED03)))
    h = _mix(s ^ key)
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (2.0 * u - 1.0) * SCALE[tensor]


class Orbitals:
    """Spin-orbital bookkeeping for (n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta)."""

    def __init__(self, noa: int, nob: int, nva: int, nvb: int):
        self.noa, self.nob, self.nva, self.nvb = noa, nob, nva, nvb
        self.Ot, self.Vt = noa + nob, nva + nvb
        self.spin_o = np.array([1] * noa + [2] * nob, dtype=np.int64)
        self.spin_v = np.array([1] * nva + [2] * nvb, dtype=np.int64)

    def orbital_energies(self) -> np.ndarray:
        """e_i = -2 + i/O (occupied), e_a = 0.5 + 2a/V (virtual); same for both spins (§8d)."""
        eo = np.concatenate([-2.0 + np.arange(self.noa) / max(self.noa, 1),
                             -2.0 + np.arange(self.nob) / max(self.nob, 1)])
        ev = np.concatenate([0.5 + 2.0 * np.arange(self.nva) / max(self.nva, 1),
                             0.5 + 2.0 * np.arange(self.nvb) / max(self.nvb, 1)])
        return np.concatenate([eo, ev])


def _antisym_pair(x, y):
    """canonical (min,max), sign (-1 if swapped), zero mask if equal."""
    lo, hi = np.minimum(x, y), np.maximum(x, y)
    sign = np.where(x > y, -1.0, 1.0)
    return lo, hi, sign, (x == y)


def block(orb: Orbitals, seed: int, tensor: int, idx) -> np.ndarray:
    """Values of `tensor` on the outer product of the index vectors idx = (i0[,i1,i2,i3])."""
    grids = np.meshgrid(*[np.asarray(i, dtype=np.int64) for i in idx], indexing="ij")
    so, sv = orb.spin_o, orb.spin_v
    if tensor == T1:
        a, i = grids
        val = raw_value(seed, T1, a, i, 0, 0)
        return np.where(sv[a] == so[i], val, 0.0)
    if tensor == T2:
        a, b, i, j = grids
        alo, ahi, s1, z1 = _antisym_pair(a, b)
        ilo, ihi, s2, z2 = _antisym_pair(i, j)
        val = raw_value(seed, T2, alo, ahi, ilo, ihi) * s1 * s2
        ok = (sv[a] + sv[b] == so[i] + so[j]) & ~z1 & ~z2
        return np.where(ok, val, 0.0)
    if tensor == V_IJAB:
        i, j, a, b = grids
        ilo, ihi, s1, z1 = _antisym_pair(i, j)
        alo, ahi, s2, z2 = _antisym_pair(a, b)
        val = raw_value(seed, V_IJAB, ilo, ihi, alo, ahi) * s1 * s2
        ok = (so[i] + so[j] == sv[a] + sv[b]) & ~z1 & ~z2
        return np.where(ok, val, 0.0)
    if tensor == V_IJKA:
        i, j, k, a = grids
        ilo, ihi, s1, z1 = _antisym_pair(i, j)
        val = raw_value(seed, V_IJKA, ilo, ihi, k, a) * s1
        ok = (so[i] + so[j] == so[k] + sv[a]) & ~z1
        return np.where(ok, val, 0.0)
    if tensor == V_IABC:
        i, a, b, c = grids
        blo, bhi, s1, z1 = _antisym_pair(b, c)
        val = raw_value(seed, V_IABC, i, a, blo, bhi) * s1
        ok = (so[i] + sv[a] == sv[b] + sv[c]) & ~z1
        return np.where(ok, val, 0.0)
    raise ValueError(tensor)


def dense(orb: Orbitals, seed: int, tensor: int) -> np.ndarray:
    """Whole tensor as a dense C-contiguous float64 array (small problems only)."""
    O, V = np.arange(orb.Ot), np.arange(orb.Vt)
    idx = {T1: (V, O), T2: (V, V, O, O), V_IJAB: (O, O, V, V), V_IJKA: (O, O, O, V),
           V_IABC: (O, V, V, V)}[tensor]
    return np.ascontiguousarray(block(orb, seed, tensor, idx), dtype=np.float64)


def dense_all(orb: Orbitals, seed: int = 1234) -> dict:
    return {"t1": dense(orb, seed, T1), "t2": dense(orb, seed, T2), "v2ijab": dense(orb, seed, V_IJAB),
            "v2ijka": dense(orb, seed, V_IJKA), "v2iabc": dense(orb, seed, V_IABC),
            "evl": orb.orbital_energies()}
