"""TEST-INPUT PROVIDER (not product code, never imported by exachem_b200/).

Closed-shell RHF -> MO integrals -> spin-orbital CCSD for small molecules, so that the (T) path can be run on
REAL converged amplitudes of the reference's own inputs (inputs/h2o.json, inputs/ci/butanol2_pt.json) and
compared with the reference's CI goldens:
  * H2O cc-pVDZ   : SCF -75.82509922164868 Eh, CCSD correlation -0.25498209984722586 Eh
                    (ci/reference_output/h2o_eom.cc-pvdz.eom_ccsd.json; same geometry/basis as inputs/h2o.json)
  * butanol2 STO-3G: SCF, CCSD and the [T]/(T) corrections of ci/reference_output/butanol2_pt.sto-3g.ccsd_t.json

Everything is written from the textbook equations: integrals by McMurchie-Davidson (gints.c), RHF with DIIS,
the spin-orbital CCSD equations of Stanton, Gauss, Watts, Bartlett (J. Chem. Phys. 94, 4334 (1991)).  The
reference builds its tensors with libint + TAMM (absent here); conventions it fixes and we follow:
  tile/spin order | occ a | occ b | virt a | virt b |         exachem/cholesky/cholesky_2e.cpp:192-226
  v2[p,q,r,s] = (pr|qs) - (ps|qr) = <pq||rs>                   exachem/cholesky/v2tensors.cpp:68-86
  d_t1[a,i], d_t2[a,b,i,j] antisymmetrised spin-orbital        exachem/cc/ccsd/cd_ccsd_cs_ann.cpp:940-975
"""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libgints.so")
ANGSTROM = 1.8897261259077822   # the reference's ang2bohr (exachem/common/constants.hpp:13)
Z_OF = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10}
L_OF = {"S": 0, "P": 1, "D": 2, "F": 3}


def _lib():
    src = os.path.join(HERE, "gints.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-fvisibility=hidden", src, "-o", SO, "-lm"])
    return C.CDLL(SO)


# ------------------------------------------------------------------------------------------------ input
def read_geometry(input_json):
    d = json.load(open(input_json))
    g = d["geometry"]
    scale = ANGSTROM if g.get("units", "angstrom").lower().startswith("ang") else 1.0
    atoms = []
    for ln in g["coordinates"]:
        f = ln.split()
        atoms.append((f[0], np.array([float(x) for x in f[1:4]]) * scale))
    return atoms, d


def read_g94(path, elements):
    """Gaussian-94 basis file -> {element: [(l, exps, coefs), ...]}; SP shells are split."""
    txt = open(path).read().replace("D+", "E+").replace("D-", "E-")
    out = {}
    for blk in txt.split("****"):
        lines = [ln for ln in blk.strip().splitlines() if ln.strip() and not ln.startswith("!")]
        if not lines:
            continue
        el = lines[0].split()[0]
        if el not in elements:
            continue
        shells, i = [], 1
        while i < len(lines):
            kind, n = lines[i].split()[0].upper(), int(lines[i].split()[1])
            rows = np.array([[float(x) for x in re.split(r"\s+", lines[i + 1 + k].strip())] for k in range(n)])
            if kind == "SP":
                shells.append((0, rows[:, 0], rows[:, 1]))
                shells.append((1, rows[:, 0], rows[:, 2]))
            else:
                shells.append((L_OF[kind], rows[:, 0], rows[:, 1]))
            i += 1 + n
        out[el] = shells
    return out


def _cart_powers(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def _sph_from_cart(l):
    """Real solid harmonics as combinations of the cartesian monomials of _cart_powers(l) (columns = pure
    functions).  Scale factors are irrelevant: every function is normalised by its self-overlap later."""
    pw = _cart_powers(l)
    idx = {p: i for i, p in enumerate(pw)}
    if l < 2:
        return np.eye(len(pw))
    if l == 2:
        M = np.zeros((6, 5))
        M[idx[(0, 0, 2)], 0], M[idx[(2, 0, 0)], 0], M[idx[(0, 2, 0)], 0] = 2.0, -1.0, -1.0   # 2zz-xx-yy
        M[idx[(1, 0, 1)], 1] = 1.0                                                            # xz
        M[idx[(0, 1, 1)], 2] = 1.0                                                            # yz
        M[idx[(2, 0, 0)], 3], M[idx[(0, 2, 0)], 3] = 1.0, -1.0                                # xx-yy
        M[idx[(1, 1, 0)], 4] = 1.0                                                            # xy
        return M
    if l == 3:
        M = np.zeros((10, 7))
        def put(col, terms):
            for pw_, c in terms:
                M[idx[pw_], col] = c
        put(0, [((0, 0, 3), 2.0), ((2, 0, 1), -3.0), ((0, 2, 1), -3.0)])          # z(2zz-3xx-3yy)
        put(1, [((1, 0, 2), 4.0), ((3, 0, 0), -1.0), ((1, 2, 0), -1.0)])          # x(4zz-xx-yy)
        put(2, [((0, 1, 2), 4.0), ((2, 1, 0), -1.0), ((0, 3, 0), -1.0)])          # y(4zz-xx-yy)
        put(3, [((2, 0, 1), 1.0), ((0, 2, 1), -1.0)])                             # z(xx-yy)
        put(4, [((1, 1, 1), 1.0)])                                                # xyz
        put(5, [((3, 0, 0), 1.0), ((1, 2, 0), -3.0)])                             # x(xx-3yy)
        put(6, [((2, 1, 0), 3.0), ((0, 3, 0), -1.0)])                             # y(3xx-yy)
        return M
    raise NotImplementedError("pure functions beyond f are not needed for the reference's small inputs")


class Molecule:
    def __init__(self, input_json, basis_dir, basisset=None):
        self.atoms, self.input = read_geometry(input_json)
        name = (basisset or self.input["basis"]["basisset"]).lower()
        bas = read_g94(os.path.join(basis_dir, name + ".g94"), {a for a, _ in self.atoms})
        centre, l, nprim, poff, coff, exps, coefs, blocks = [], [], [], [], [], [], [], []
        nc = 0
        for el, xyz in self.atoms:
            for (ll, ex, co) in bas[el]:
                centre.append(xyz); l.append(ll); nprim.append(len(ex)); poff.append(len(exps)); coff.append(nc)
                exps.extend(ex)
                coefs.extend(co * ex ** ((2 * ll + 3) / 4.0))     # contraction over normalised primitives
                blocks.append((nc, _sph_from_cart(ll)))
                nc += len(_cart_powers(ll))
        self.ncart = nc
        nsph = sum(b.shape[1] for _, b in blocks)
        U = np.zeros((nc, nsph))
        c0 = 0
        for off, b in blocks:
            U[off:off + b.shape[0], c0:c0 + b.shape[1]] = b
            c0 += b.shape[1]
        self.U = U
        self.sh = dict(centre=np.ascontiguousarray(centre, np.float64).ravel(), l=np.array(l, np.int32),
                       nprim=np.array(nprim, np.int32), poff=np.array(poff, np.int32), coff=np.array(coff, np.int32),
                       exps=np.array(exps, np.float64), coefs=np.array(coefs, np.float64))
        self.Z = np.array([Z_OF[a] for a, _ in self.atoms], np.float64)
        self.xyz = np.ascontiguousarray([x for _, x in self.atoms], np.float64).ravel()
        self.nelec = int(self.Z.sum()) - int(self.input.get("SCF", {}).get("charge", 0))

    def nuclear_repulsion(self):
        e, X = 0.0, self.xyz.reshape(-1, 3)
        for i in range(len(self.Z)):
            for j in range(i):
                e += self.Z[i] * self.Z[j] / np.linalg.norm(X[i] - X[j])
        return e

    def integrals(self):
        """S, T, V (n x n) and (pq|rs) (n^4) over normalised real spherical functions."""
        L = _lib()
        s, n = self.sh, self.ncart
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
        S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        L.gints_one_electron(len(s["l"]), p(s["centre"], dp), p(s["l"], ip), p(s["nprim"], ip), p(s["poff"], ip),
                             p(s["coff"], ip), p(s["exps"], dp), p(s["coefs"], dp), len(self.Z), p(self.Z, dp),
                             p(self.xyz, dp), n, p(S, dp), p(T, dp), p(V, dp))
        eri = np.zeros((n, n, n, n))
        L.gints_eri(len(s["l"]), p(s["centre"], dp), p(s["l"], ip), p(s["nprim"], ip), p(s["poff"], ip),
                    p(s["coff"], ip), p(s["exps"], dp), p(s["coefs"], dp), n, p(eri, dp))
        U = self.U
        S, T, V = U.T @ S @ U, U.T @ T @ U, U.T @ V @ U
        for ax in range(4):
            eri = np.moveaxis(np.tensordot(eri, U, axes=([ax], [0])), -1, ax)
        nrm = 1.0 / np.sqrt(np.diag(S))
        S, T, V = S * np.outer(nrm, nrm), T * np.outer(nrm, nrm), V * np.outer(nrm, nrm)
        eri = eri * nrm[:, None, None, None] * nrm[None, :, None, None] * nrm[None, None, :, None] * nrm[None, None, None, :]
        return S, T, V, np.ascontiguousarray(eri)


# ------------------------------------------------------------------------------------------------ RHF
def rhf(S, H, eri, nocc, enuc, conv=1e-12, maxiter=200):
    s, Us = np.linalg.eigh(S)
    X = Us @ np.diag(s ** -0.5) @ Us.T
    def diag(F):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        return e, X @ Cp
    e, Cm = diag(H)
    Fs, Es, Eold = [], [], 0.0
    for it in range(maxiter):
        D = Cm[:, :nocc] @ Cm[:, :nocc].T
        J = np.einsum("pqrs,rs->pq", eri, D)
        K = np.einsum("prqs,rs->pq", eri, D)
        F = H + 2 * J - K
        E = np.sum(D * (H + F)) + enuc
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-10:], Es[-10:]
        if len(Fs) > 1:
            m = len(Fs)
            B = -np.ones((m + 1, m + 1)); B[m, m] = 0
            for i in range(m):
                for j in range(m):
                    B[i, j] = np.sum(Es[i] * Es[j])
            rhs = np.zeros(m + 1); rhs[m] = -1
            c = np.linalg.solve(B, rhs)[:m]
            F = sum(ci * Fi for ci, Fi in zip(c, Fs))
        e, Cm = diag(F)
        if abs(E - Eold) < conv and np.abs(err).max() < 1e-9:
            break
        Eold = E
    return E, e, Cm, it + 1


def uhf(S, H, eri, na, nb, enuc, conv=1e-12, maxiter=300):
    """unrestricted Hartree-Fock with DIIS; returns E, (eps_a, eps_b), (C_a, C_b), iterations"""
    s, Us = np.linalg.eigh(S)
    X = Us @ np.diag(s ** -0.5) @ Us.T
    def diag(F):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        return e, X @ Cp
    ea, Ca = diag(H)
    eb, Cb = ea.copy(), Ca.copy()
    if na != nb:   # break the alpha/beta symmetry of the core guess a little
        Cb = Cb.copy(); Cb[:, [nb - 1, nb]] = Cb[:, [nb, nb - 1]]
    hist_f, hist_e, Eold = [], [], 0.0
    for it in range(maxiter):
        Da, Db = Ca[:, :na] @ Ca[:, :na].T, Cb[:, :nb] @ Cb[:, :nb].T
        J = np.einsum("pqrs,rs->pq", eri, Da + Db)
        Fa = H + J - np.einsum("prqs,rs->pq", eri, Da)
        Fb = H + J - np.einsum("prqs,rs->pq", eri, Db)
        E = 0.5 * (np.sum((Da + Db) * H) + np.sum(Da * Fa) + np.sum(Db * Fb)) + enuc
        err = np.concatenate([(X.T @ (Fa @ Da @ S - S @ Da @ Fa) @ X).ravel(), (X.T @ (Fb @ Db @ S - S @ Db @ Fb) @ X).ravel()])
        hist_f.append((Fa, Fb)); hist_e.append(err)
        hist_f, hist_e = hist_f[-10:], hist_e[-10:]
        if len(hist_f) > 1:
            m = len(hist_f)
            B = -np.ones((m + 1, m + 1)); B[m, m] = 0
            for i in range(m):
                for j in range(m):
                    B[i, j] = hist_e[i] @ hist_e[j]
            rhs = np.zeros(m + 1); rhs[m] = -1
            c = np.linalg.solve(B, rhs)[:m]
            Fa = sum(ci * f[0] for ci, f in zip(c, hist_f))
            Fb = sum(ci * f[1] for ci, f in zip(c, hist_f))
        ea, Ca = diag(Fa)
        eb, Cb = diag(Fb)
        if abs(E - Eold) < conv and np.abs(err).max() < 1e-9:
            break
        Eold = E
    return E, (ea, eb), (Ca, Cb), it + 1


def solve_uhf(input_json, basis_dir, multiplicity, basisset=None, verbose=True):
    """open-shell route: integrals -> UHF -> spin-orbital CCSD.  Returns the dense spin-orbital tensors in the reference's
    layout (| occ a | occ b | virt a | virt b |) and a summary; [T]/(T) need is_restricted = false."""
    mol = Molecule(input_json, basis_dir, basisset)
    S, T, V, eri = mol.integrals()
    nb = (mol.nelec - (multiplicity - 1)) // 2
    na = mol.nelec - nb
    escf, (ea, eb), (Ca, Cb), it_scf = uhf(S, T + V, eri, na, nb, mol.nuclear_repulsion())
    n = len(ea)
    Cso = np.concatenate([Ca[:, :na], Cb[:, :nb], Ca[:, na:], Cb[:, nb:]], axis=1)       # spatial parts, tile order
    spin = np.array([0] * na + [1] * nb + [0] * (n - na) + [1] * (n - nb))
    eso = np.concatenate([ea[:na], eb[:nb], ea[na:], eb[nb:]])
    g = np.einsum("pqrs,pi,qj,rk,sl->ijkl", eri, Cso, Cso, Cso, Cso, optimize=True)       # (ij|kl) over spin orbitals
    same = (spin[:, None] == spin[None, :]).astype(float)
    g = g * same[:, :, None, None] * same[None, None, :, :]
    phys = g.transpose(0, 2, 1, 3)
    anti = phys - phys.transpose(0, 1, 3, 2)
    no = na + nb
    ecc, t1, t2, it_cc = ccsd(anti, eso, no, verbose=False)
    o, v = slice(0, no), slice(no, None)
    c = np.ascontiguousarray
    tensors = dict(evl=eso, t1=c(t1.T), t2=c(t2.transpose(2, 3, 0, 1)), v2ijab=c(anti[o, o, v, v]),
                   v2ijka=c(anti[o, o, o, v]), v2iabc=c(anti[o, v, v, v]))
    s2 = 0.25 * (na - nb) * (na - nb + 2) + nb - np.sum((Ca[:, :na].T @ S @ Cb[:, :nb]) ** 2)
    info = dict(nbf=int(n), n_occ_alpha=int(na), n_occ_beta=int(nb), e_nuc=float(mol.nuclear_repulsion()), e_scf=float(escf),
                e_ccsd_corr=float(ecc), s_squared=float(s2), scf_iterations=int(it_scf), ccsd_iterations=int(it_cc))
    if verbose:
        print(info)
    return tensors, info


# ------------------------------------------------------------------------------------------------ CD
def pivoted_cholesky(eri_mo, diagtol):
    """Pivoted incomplete Cholesky of the (pq|rs) matrix to diagonal tolerance `diagtol` (the approximation
    the reference applies before CCSD: CD.diagtol, exachem/cholesky/cholesky_2e.cpp); returns the
    reconstructed integrals and the number of vectors."""
    n = eri_mo.shape[0]
    M = eri_mo.reshape(n * n, n * n)
    d = np.diag(M).copy()
    Lv = []
    while True:
        piv = int(np.argmax(d))
        if d[piv] <= diagtol:
            break
        col = M[:, piv].copy()
        for v in Lv:
            col -= v * v[piv]
        v = col / np.sqrt(d[piv])
        Lv.append(v)
        d -= v * v
    Lm = np.array(Lv)
    return (Lm.T @ Lm).reshape(n, n, n, n), len(Lv)


# ------------------------------------------------------------------------------------------------ CCSD
class SpinOrbitalIntegrals:
    """<pq||rs> over spin orbitals ordered | occ a | occ b | virt a | virt b |, built block by block from the
    spatial (pq|rs) (the full spin-orbital tensor of benzene would be 21 GB)."""

    def __init__(self, eps, eri_mo, nocc):
        n = len(eps)
        nvir = n - nocc
        self.eri = eri_mo
        self.spat = np.concatenate([np.arange(nocc), np.arange(nocc), np.arange(nocc, n), np.arange(nocc, n)])
        self.spin = np.concatenate([np.zeros(nocc, int), np.ones(nocc, int), np.zeros(nvir, int), np.ones(nvir, int)])
        self.eso = eps[self.spat]
        self.no = 2 * nocc
        self.rng = {"o": np.arange(0, 2 * nocc), "v": np.arange(2 * nocc, 2 * n)}

    def block(self, kinds):
        """kinds e.g. "oovv" -> <ij||ab>"""
        P, Q, R, S = (self.rng[k] for k in kinds)
        sp, sn = self.spat, self.spin
        d = lambda a, b: (sn[a][:, None] == sn[b][None, :]).astype(float)  # noqa: E731
        # <pq|rs> = (pr|qs)
        out = self.eri[np.ix_(sp[P], sp[R], sp[Q], sp[S])].transpose(0, 2, 1, 3) * d(P, R)[:, None, :, None] * d(Q, S)[None, :, None, :]
        out = np.ascontiguousarray(out)
        x = self.eri[np.ix_(sp[P], sp[S], sp[Q], sp[R])].transpose(0, 2, 3, 1) * d(P, S)[:, None, None, :] * d(Q, R)[None, :, :, None]
        out -= x
        return out


def spin_orbital_integrals(eps, eri_mo, nocc):
    """full <pq||rs> tensor (small molecules only), orbital energies and spins"""
    so = SpinOrbitalIntegrals(eps, eri_mo, nocc)
    allr = np.arange(len(so.spat))
    so.rng = {"o": allr, "v": allr}
    return so.block("oooo"), so.eso, so.spin


def ccsd(ints, eso=None, no=None, conv=1e-11, maxiter=200, ndiis=8, verbose=False):
    """Stanton et al. spin-orbital CCSD (canonical RHF orbitals: f is diagonal).  `ints` is a SpinOrbitalIntegrals
    (or, with eso and no, a full <pq||rs> array).  W_abef is never stored: its three pieces are contracted with
    tau directly.  Returns E_corr, t1[i,a], t2[i,j,a,b], iterations."""
    es = lambda *a: np.einsum(*a, optimize=True)  # noqa: E731
    if isinstance(ints, SpinOrbitalIntegrals):
        eso, no = ints.eso, ints.no
        blk = ints.block
    else:
        o_, v_ = slice(0, no), slice(no, None)
        sl = {"o": o_, "v": v_}
        blk = lambda k: ints[sl[k[0]], sl[k[1]], sl[k[2]], sl[k[3]]]  # noqa: E731
    nv = len(eso) - no
    fo, fv = eso[:no], eso[no:]
    Dia = fo[:, None] - fv[None, :]
    Dijab = fo[:, None, None, None] + fo[None, :, None, None] - fv[None, None, :, None] - fv[None, None, None, :]
    oovv, ooov, ovvv = blk("oovv"), blk("ooov"), blk("ovvv")
    oooo, vvvv, ovvo, ovov = blk("oooo"), blk("vvvv"), blk("ovvo"), blk("ovov")
    vvvv2 = vvvv.reshape(nv * nv, nv * nv)
    t1 = np.zeros((no, nv))
    t2 = oovv / Dijab
    E = 0.25 * np.sum(oovv * t2)
    hist_t, hist_e = [], []
    for it in range(maxiter):
        tt = es("ia,jb->ijab", t1, t1)
        ttau = t2 + 0.5 * (tt - tt.transpose(0, 1, 3, 2))
        tau = t2 + tt - tt.transpose(0, 1, 3, 2)
        del tt
        Fae = es("mf,mafe->ae", t1, ovvv) - 0.5 * es("mnaf,mnef->ae", ttau, oovv)
        Fmi = es("ne,mnie->mi", t1, ooov) + 0.5 * es("inef,mnef->mi", ttau, oovv)
        Fme = es("nf,mnef->me", t1, oovv)
        Wmnij = oooo + es("je,mnie->mnij", t1, ooov) - es("ie,mnje->mnij", t1, ooov) + 0.25 * es("ijef,mnef->mnij", tau, oovv)
        # <mn||ej> = -<mn||je>
        Wmbej = ovvo + es("jf,mbef->mbej", t1, ovvv) + es("nb,mnje->mbej", t1, ooov) \
            - es("jnfb,mnef->mbej", 0.5 * t2 + es("jf,nb->jnfb", t1, t1), oovv)
        # T1
        r1 = (es("ie,ae->ia", t1, Fae) - es("ma,mi->ia", t1, Fmi) + es("imae,me->ia", t2, Fme)
              - es("nf,naif->ia", t1, ovov) - 0.5 * es("imef,maef->ia", t2, ovvv)
              + 0.5 * es("mnae,nmie->ia", t2, ooov))          # -1/2 t_mn^ae <nm||ei>, <nm||ei> = -<nm||ie>
        # T2
        Fbe = Fae - 0.5 * es("mb,me->be", t1, Fme)
        Fmj = Fmi + 0.5 * es("je,me->mj", t1, Fme)
        x = es("ijae,be->ijab", t2, Fbe)
        r2 = oovv + x - x.transpose(0, 1, 3, 2)
        x = es("imab,mj->ijab", t2, Fmj)
        r2 -= x - x.transpose(1, 0, 2, 3)
        r2 += 0.5 * es("mnab,mnij->ijab", tau, Wmnij)
        # 1/2 tau_ij^ef W_abef with W_abef = <ab||ef> + P(ab) t_m^b <ma||ef> + 1/4 tau_mn^ab <mn||ef>
        r2 += 0.5 * (tau.reshape(no * no, nv * nv) @ vvvv2.T).reshape(no, no, nv, nv)
        y = es("ijef,maef->ijma", tau, ovvv)                       # sum_ef tau_ij^ef <ma||ef>
        x = es("mb,ijma->ijab", t1, y)
        r2 += 0.5 * (x - x.transpose(0, 1, 3, 2))
        y = es("ijef,mnef->ijmn", tau, oovv)
        r2 += 0.125 * es("mnab,ijmn->ijab", tau, y)
        x = es("imae,mbej->ijab", t2, Wmbej) - es("ie,ma,mbej->ijab", t1, t1, ovvo)
        r2 += x - x.transpose(1, 0, 2, 3) - x.transpose(0, 1, 3, 2) + x.transpose(1, 0, 3, 2)
        # P(ij) t_i^e <ab||ej>, <ab||ej> = <ej||ab> = -<je||ab>... use ovvv: <ab||ej> = <ej||ab> = -<je||ab> = <je||ba>
        x = es("ie,jeba->ijab", t1, ovvv)
        r2 += x - x.transpose(1, 0, 2, 3)
        # -P(ab) t_m^a <mb||ij>, <mb||ij> = <ij||mb> = ooov[i,j,m,b]
        x = es("ma,ijmb->ijab", t1, ooov)
        r2 -= x - x.transpose(0, 1, 3, 2)
        n1, n2 = r1 / Dia, r2 / Dijab
        res = np.sqrt(np.sum((n1 - t1) ** 2) + np.sum((n2 - t2) ** 2))
        vec = np.concatenate([n1.ravel(), n2.ravel()])
        hist_t.append(vec); hist_e.append(vec - np.concatenate([t1.ravel(), t2.ravel()]))
        hist_t, hist_e = hist_t[-ndiis:], hist_e[-ndiis:]
        if len(hist_t) > 1:
            m = len(hist_t)
            B = -np.ones((m + 1, m + 1)); B[m, m] = 0
            for i in range(m):
                for j in range(m):
                    B[i, j] = hist_e[i] @ hist_e[j]
            rhs = np.zeros(m + 1); rhs[m] = -1
            c = np.linalg.solve(B, rhs)[:m]
            vec = sum(ci * ti for ci, ti in zip(c, hist_t))
        t1, t2 = vec[:no * nv].reshape(no, nv), vec[no * nv:].reshape(no, no, nv, nv)
        Enew = 0.25 * np.sum(oovv * t2) + 0.5 * es("ijab,ia,jb->", oovv, t1, t1)
        if verbose:
            print(f"  ccsd iter {it + 1}: E = {Enew:.12f}  residual {res:.3e}", flush=True)
        if res < conv and abs(Enew - E) < conv:
            E = Enew
            break
        E = Enew
    return E, t1, t2, it + 1


# ------------------------------------------------------------------------------------------------ fixtures
def closed_shell_fixture(eps, eri_mo, nocc, t1_so, t2_so):
    """Compact (spatial-orbital) form of everything the (T) path consumes: orbital energies, the (ov|ov),
    (oo|ov), (ov|vv) MO integrals and the alpha / alpha-beta CCSD amplitudes."""
    n = len(eps)
    nv = n - nocc
    o, v = slice(0, nocc), slice(nocc, n)
    return dict(eps=eps, nocc=np.int64(nocc),
                ovov=np.ascontiguousarray(eri_mo[o, v, o, v]), ooov=np.ascontiguousarray(eri_mo[o, o, o, v]),
                ovvv=np.ascontiguousarray(eri_mo[o, v, v, v]),
                t1=np.ascontiguousarray(t1_so[:nocc, :nv]),                                # t1[I,A], alpha
                t2=np.ascontiguousarray(t2_so[:nocc, nocc:2 * nocc, :nv, nv:2 * nv]))      # t2[I,J,A,B], i,a alpha; j,b beta


def spin_orbital_tensors(fx):
    """Fixture -> the five dense spin-orbital tensors + orbital energies in the reference's layout
    (| occ a | occ b | virt a | virt b |; T1[a,i], T2[a,b,i,j], v2ijab[i,j,a,b], v2ijka[i,j,k,a], v2iabc[i,a,b,c])."""
    eps, no = np.asarray(fx["eps"]), int(fx["nocc"])
    nv = len(eps) - no
    io = np.concatenate([np.arange(no), np.arange(no)]); so = np.concatenate([np.zeros(no, int), np.ones(no, int)])
    iv = np.concatenate([np.arange(nv), np.arange(nv)]); sv = np.concatenate([np.zeros(nv, int), np.ones(nv, int)])
    d = lambda a, b: (a[:, None] == b[None, :]).astype(float)  # noqa: E731
    oo, ov, vv = d(so, so), d(so, sv), d(sv, sv)
    ovov, ooov, ovvv = fx["ovov"], fx["ooov"], fx["ovvv"]
    # <ij||ab> = (ia|jb) - (ib|ja)
    g = ovov[np.ix_(io, iv, io, iv)]                                                    # [i,a,j,b]
    v2ijab = g.transpose(0, 2, 1, 3) * ov[:, None, :, None] * ov[None, :, None, :] \
        - g.transpose(0, 2, 3, 1) * ov[:, None, None, :] * ov[None, :, :, None]
    # <ij||ka> = (ik|ja) - (jk|ia)
    g = ooov[np.ix_(io, io, io, iv)]                                                    # [p,q,r,a] = (pq|ra)
    v2ijka = g.transpose(0, 2, 1, 3) * oo[:, None, :, None] * ov[None, :, None, :] \
        - g.transpose(2, 0, 1, 3) * oo[None, :, :, None] * ov[:, None, None, :]
    # <ia||bc> = (ib|ac) - (ic|ab)
    g = ovvv[np.ix_(io, iv, iv, iv)]                                                    # [i,p,q,r] = (ip|qr)
    v2iabc = g.transpose(0, 2, 1, 3) * ov[:, None, :, None] * vv[None, :, None, :] \
        - g.transpose(0, 2, 3, 1) * ov[:, None, None, :] * vv[None, :, :, None]
    t = fx["t2"][np.ix_(io, io, iv, iv)]                                                # [i,j,a,b] spatial parts
    t2 = t * ov[:, None, :, None] * ov[None, :, None, :] - t.transpose(0, 1, 3, 2) * ov[:, None, None, :] * ov[None, :, :, None]
    t1 = fx["t1"][np.ix_(io, iv)] * ov
    c = np.ascontiguousarray
    return dict(evl=np.concatenate([eps[:no], eps[:no], eps[no:], eps[no:]]), t1=c(t1.T), t2=c(t2.transpose(2, 3, 0, 1)),
                v2ijab=c(v2ijab), v2ijka=c(v2ijka), v2iabc=c(v2iabc))


def solve(input_json, basis_dir, diagtol=None, verbose=True):
    """integrals -> RHF -> (optional Cholesky truncation) -> CCSD.  Returns (fixture dict, summary dict)."""
    mol = Molecule(input_json, basis_dir)
    S, T, V, eri = mol.integrals()
    no = mol.nelec // 2
    escf, eps, Cm, it_scf = rhf(S, T + V, eri, no, mol.nuclear_repulsion())
    eri_mo = np.einsum("pqrs,pi,qj,rk,sl->ijkl", eri, Cm, Cm, Cm, Cm, optimize=True)
    ncv = None
    if diagtol:
        eri_mo, ncv = pivoted_cholesky(eri_mo, diagtol)
    ecc, t1, t2, it_cc = ccsd(SpinOrbitalIntegrals(eps, eri_mo, no), verbose=verbose)
    info = dict(nbf=int(len(eps)), nocc=int(no), e_nuc=float(mol.nuclear_repulsion()), e_scf=float(escf),
                e_ccsd_corr=float(ecc), scf_iterations=int(it_scf), ccsd_iterations=int(it_cc), cholesky_vectors=ncv)
    if verbose:
        print(info)
    return closed_shell_fixture(eps, eri_mo, no, t1, t2), info
