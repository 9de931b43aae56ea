"""TEST INFRASTRUCTURE (not shipped, never imported by ``auto_oo_b200``): a small Gaussian-integral
and RHF code, so that the reference's *molecular* golden values can be replayed without PySCF.

The reference obtains its inputs from PySCF (``moldata_pyscf.py:28-35``: ``int1e_kin + int1e_nuc``,
``int2e``, ``int1e_ovlp``, ``get_enuc`` and ``mol.RHF().run()``), which is not installed here.  This
module restates the published algorithms that produce the same numbers for the STO-3G basis:

* McMurchie-Davidson Hermite-Gaussian scheme for overlap, kinetic, nuclear-attraction and
  electron-repulsion integrals over contracted Cartesian Gaussians (McMurchie & Davidson,
  J. Comput. Phys. 26, 218 (1978); Helgaker, Jorgensen, Olsen, "Molecular Electronic-Structure
  Theory", ch. 9), Boys function through the confluent hypergeometric function;
* STO-3G exponents / contraction coefficients (Hehre, Stewart, Pople, J. Chem. Phys. 51, 2657
  (1969)) for H, C, N, O in the digits the EMSL basis-set exchange distributes;
* PySCF conventions that the golden vectors depend on: Z-matrix to Cartesian conversion
  (``pyscf.gto.mole.from_zmatrix``), Angstrom input with ``BOHR = 0.52917721092``, AO order
  atom-by-atom ``1s, 2s, 2px, 2py, 2pz``, contracted functions normalised to one;
* restricted Hartree-Fock with DIIS.

It is pinned by ``tests/test_molecular_golden.py`` against the golden numbers printed in the
reference's own tests (``test/test_oo_energy.py:27-103`` RHF orbitals in the OAO basis, ``:240-312``
energy at given orbitals / RDMs, ``:317-412`` orbital-optimised energy).
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.special import hyp1f1

BOHR = 0.52917721092            # pyscf.data.nist.BOHR (Angstrom)

_S_COEF = (0.1543289673, 0.5353281423, 0.4446345422)
_SP_S_COEF = (-0.09996722919, 0.3995128261, 0.7001154689)
_SP_P_COEF = (0.1559162750, 0.6076837186, 0.3919573931)
STO3G = {
    "H": {"Z": 1, "shells": [(0, (3.425250914, 0.6239137298, 0.1688554040), _S_COEF)]},
    "C": {"Z": 6, "shells": [(0, (71.61683735, 13.04509632, 3.530512160), _S_COEF),
                             (0, (2.941249355, 0.6834830964, 0.2222899159), _SP_S_COEF),
                             (1, (2.941249355, 0.6834830964, 0.2222899159), _SP_P_COEF)]},
    "N": {"Z": 7, "shells": [(0, (99.10616896, 18.05231239, 4.885660238), _S_COEF),
                             (0, (3.780455879, 0.8784966449, 0.2857143744), _SP_S_COEF),
                             (1, (3.780455879, 0.8784966449, 0.2857143744), _SP_P_COEF)]},
    "O": {"Z": 8, "shells": [(0, (130.7093214, 23.80886605, 6.443608313), _S_COEF),
                             (0, (5.033151319, 1.169596125, 0.3803889600), _SP_S_COEF),
                             (1, (5.033151319, 1.169596125, 0.3803889600), _SP_P_COEF)]},
}


# ------------------------------------------------------------------------------------------
# geometry
def _rotation_mat(vec, theta):
    """Right-handed rotation by ``theta`` about ``vec`` (Rodrigues)."""
    vec = np.asarray(vec, dtype=float)
    vec = vec / np.linalg.norm(vec)
    uu = np.outer(vec, vec)
    ux = np.array([[0, -vec[2], vec[1]], [vec[2], 0, -vec[0]], [-vec[1], vec[0], 0]])
    c, s = math.cos(theta), math.sin(theta)
    return c * np.eye(3) + s * ux + (1 - c) * uu


def from_zmatrix(text):
    """Z-matrix (Angstrom / degrees) -> list of (symbol, xyz in Angstrom), following the placement
    rules of ``pyscf.gto.mole.from_zmatrix``: atom 1 at the origin, atom 2 on +x, atom 3 rotated about
    ``v1 x z``, further atoms by bond / angle / dihedral."""
    symb, coord = [], []
    for line in text.replace(";", "\n").replace(",", " ").splitlines():
        raw = line.split()
        if not raw or raw[0].startswith("#"):
            continue
        symb.append(raw[0])
        if len(raw) < 3:
            coord.append(np.zeros(3))
        elif len(raw) == 3:
            coord.append(np.array([float(raw[2]), 0.0, 0.0]))
        elif len(raw) == 5:
            bonda, bond, anga, ang = int(raw[1]) - 1, float(raw[2]), int(raw[3]) - 1, float(raw[4]) / 180 * np.pi
            v1 = coord[anga] - coord[bonda]
            vecn = np.cross(v1, [0.0, 0.0, 1.0]) if not np.allclose(v1[:2], 0) else np.array([0.0, 0.0, 1.0])
            c = _rotation_mat(vecn, ang) @ v1 * (bond / np.linalg.norm(v1))
            coord.append(coord[bonda] + c)
        else:
            bonda, bond, anga, ang = int(raw[1]) - 1, float(raw[2]), int(raw[3]) - 1, float(raw[4]) / 180 * np.pi
            v1 = coord[anga] - coord[bonda]
            v1 = v1 / np.linalg.norm(v1)
            if ang < 1e-7:
                c = v1 * bond
            elif np.pi - ang < 1e-7:
                c = -v1 * bond
            else:
                diha, dih = int(raw[5]) - 1, float(raw[6]) / 180 * np.pi
                v2 = coord[diha] - coord[anga]
                vecn = np.cross(v2, -v1)
                nrm = np.linalg.norm(vecn)
                if nrm < 1e-7:
                    vecn = np.cross(v1, [0.0, 0.0, 1.0]) if not np.allclose(v1[:2], 0) else np.array([0., 0., 1.])
                    c = _rotation_mat(vecn, ang) @ v1 * bond
                else:
                    vecn = _rotation_mat(v1, -dih) @ vecn / nrm
                    c = _rotation_mat(vecn, ang) @ v1 * bond
            coord.append(coord[bonda] + c)
    return list(zip(symb, coord))


def formaldimine_geometry(alpha, phi):
    """The reference's ``get_formal_geo(alpha, phi)`` Z-matrix (``utils/miscellaneous.py:34-46``)."""
    return ("N\nC 1 1.498047\nH 2 1.066797 1 118.359375\nH 2 1.066797 1 118.359375 3 180\n"
            f"H 1 0.987109 2 {alpha} 3 {phi}\n")


def water_geometry(r=0.9584, angle=104.45):
    return f"O\nH 1 {r}\nH 1 {r} 2 {angle}\n"


# ------------------------------------------------------------------------------------------
# basis: one entry per contracted Cartesian function
def _dfact(n):
    return 1.0 if n <= 0 else n * _dfact(n - 2)


def _prim_norm(a, lmn):
    l, m, n = lmn
    L = l + m + n
    return ((2 * a / np.pi) ** 0.75 * (4 * a) ** (L / 2)
            / math.sqrt(_dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1)))


class Basis:
    """Contracted Cartesian Gaussians in PySCF AO order (per atom: shells in basis order, p = x,y,z)."""

    def __init__(self, atoms_bohr):
        self.centers, self.lmn, self.exps, self.coefs = [], [], [], []
        for sym, xyz in atoms_bohr:
            for l, exps, coefs in STO3G[sym]["shells"]:
                comps = [(0, 0, 0)] if l == 0 else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
                for lmn in comps:
                    e = np.asarray(exps, dtype=float)
                    c = np.asarray(coefs, dtype=float) * np.array([_prim_norm(a, lmn) for a in e])
                    self.centers.append(np.asarray(xyz, dtype=float))
                    self.lmn.append(lmn)
                    self.exps.append(e)
                    self.coefs.append(c)
        self.nao = len(self.lmn)
        # normalise every contracted function to one (PySCF does the same when it builds the env)
        for i in range(self.nao):
            s = _overlap_pair(self, i, i)
            self.coefs[i] = self.coefs[i] / math.sqrt(s)


# ------------------------------------------------------------------------------------------
# Hermite expansion coefficients E_t^{ij} for a pair of 1-D primitives (vectorised over primitive pairs)
def _hermite_E(i, j, a, b, Q):
    """dict t -> array; a, b, Q arrays over primitive pairs.  Q = A - B along this axis."""
    p = a + b
    q = a * b / p
    PA = -b * Q / p
    PB = a * Q / p
    tab = {(0, 0): {0: np.exp(-q * Q * Q)}}

    def get(ii, jj, t):
        return tab.get((ii, jj), {}).get(t, 0.0) if (ii >= 0 and jj >= 0 and 0 <= t <= ii + jj) else 0.0

    for ii in range(i + 1):
        for jj in range(j + 1):
            if ii == 0 and jj == 0:
                continue
            tab[(ii, jj)] = {}
            for t in range(ii + jj + 1):
                if ii > 0:
                    v = get(ii - 1, jj, t - 1) / (2 * p) + PA * get(ii - 1, jj, t) + (t + 1) * get(ii - 1, jj, t + 1)
                else:
                    v = get(ii, jj - 1, t - 1) / (2 * p) + PB * get(ii, jj - 1, t) + (t + 1) * get(ii, jj - 1, t + 1)
                tab[(ii, jj)][t] = v
    return tab[(i, j)]


def _pair_prims(bas, i, j):
    a = np.repeat(bas.exps[i], len(bas.exps[j]))
    b = np.tile(bas.exps[j], len(bas.exps[i]))
    c = np.outer(bas.coefs[i], bas.coefs[j]).ravel()
    return a, b, c


def _overlap_1d(i, j, a, b, Q):
    return _hermite_E(i, j, a, b, Q)[0] * np.sqrt(np.pi / (a + b))


def _overlap_pair(bas, i, j):
    a, b, c = _pair_prims(bas, i, j)
    A, B = bas.centers[i], bas.centers[j]
    s = c.copy()
    for ax in range(3):
        s = s * _overlap_1d(bas.lmn[i][ax], bas.lmn[j][ax], a, b, A[ax] - B[ax])
    return float(s.sum())


def _kinetic_pair(bas, i, j):
    a, b, c = _pair_prims(bas, i, j)
    A, B = bas.centers[i], bas.centers[j]
    li, lj = bas.lmn[i], bas.lmn[j]
    S = [[None] * 3 for _ in range(3)]          # S[ax][k] with k = j-shift index (-2, 0, +2) -> 0, 1, 2
    for ax in range(3):
        for k, sh in enumerate((-2, 0, 2)):
            jj = lj[ax] + sh
            S[ax][k] = _overlap_1d(li[ax], jj, a, b, A[ax] - B[ax]) if jj >= 0 else 0.0
    total = 0.0
    for ax in range(3):
        l = lj[ax]
        t1d = -0.5 * (l * (l - 1) * S[ax][0] - 2 * b * (2 * l + 1) * S[ax][1] + 4 * b * b * S[ax][2])
        o = [S[k][1] for k in range(3) if k != ax]
        total = total + t1d * o[0] * o[1]
    return float((c * total).sum())


def _boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2 * n + 1)


def _hermite_R(L, alpha, X, Y, Z):
    """All Hermite Coulomb integrals R_{tuv} with t+u+v <= L (arrays over the leading shape of alpha)."""
    r2 = X * X + Y * Y + Z * Z
    # R^n_{000}
    Rn = {(n, 0, 0, 0): (-2 * alpha) ** n * _boys(n, alpha * r2) for n in range(L + 1)}

    def get(n, t, u, v):
        if t < 0 or u < 0 or v < 0:
            return 0.0
        key = (n, t, u, v)
        if key not in Rn:
            if t > 0:
                val = (t - 1) * get(n + 1, t - 2, u, v) + X * get(n + 1, t - 1, u, v)
            elif u > 0:
                val = (u - 1) * get(n + 1, t, u - 2, v) + Y * get(n + 1, t, u - 1, v)
            else:
                val = (v - 1) * get(n + 1, t, u, v - 2) + Z * get(n + 1, t, u, v - 1)
            Rn[key] = val
        return Rn[key]

    return {(t, u, v): get(0, t, u, v) for t in range(L + 1) for u in range(L + 1 - t) for v in range(L + 1 - t - u)}


def _nuclear_pair(bas, i, j, charges, positions):
    a, b, c = _pair_prims(bas, i, j)
    A, B = bas.centers[i], bas.centers[j]
    li, lj = bas.lmn[i], bas.lmn[j]
    p = a + b
    P = (a[:, None] * A[None, :] + b[:, None] * B[None, :]) / p[:, None]
    E = [_hermite_E(li[ax], lj[ax], a, b, A[ax] - B[ax]) for ax in range(3)]
    L = sum(li) + sum(lj)
    total = 0.0
    for Zc, C in zip(charges, positions):
        R = _hermite_R(L, p, P[:, 0] - C[0], P[:, 1] - C[1], P[:, 2] - C[2])
        acc = 0.0
        for t, u, v in itertools.product(range(li[0] + lj[0] + 1), range(li[1] + lj[1] + 1), range(li[2] + lj[2] + 1)):
            acc = acc + E[0][t] * E[1][u] * E[2][v] * R[(t, u, v)]
        total = total - Zc * float((c * 2 * np.pi / p * acc).sum())
    return total


def _eri_all(bas):
    """(ij|kl) for all contracted functions, chemist notation, 8-fold symmetry filled in."""
    n = bas.nao
    pairs = [(i, j) for i in range(n) for j in range(i + 1)]
    # per pair: primitive-pair arrays and Hermite coefficients per (t,u,v)
    pdata = []
    for i, j in pairs:
        a, b, c = _pair_prims(bas, i, j)
        A, B = bas.centers[i], bas.centers[j]
        p = a + b
        P = (a[:, None] * A[None, :] + b[:, None] * B[None, :]) / p[:, None]
        li, lj = bas.lmn[i], bas.lmn[j]
        E = [_hermite_E(li[ax], lj[ax], a, b, A[ax] - B[ax]) for ax in range(3)]
        herm = {}
        for t, u, v in itertools.product(range(li[0] + lj[0] + 1), range(li[1] + lj[1] + 1),
                                         range(li[2] + lj[2] + 1)):
            herm[(t, u, v)] = c * E[0][t] * E[1][u] * E[2][v]
        pdata.append((p, P, herm, sum(li) + sum(lj)))
    g = np.zeros((n, n, n, n))
    for ab, (i, j) in enumerate(pairs):
        p, P, hab, Lab = pdata[ab]
        for cd in range(ab + 1):
            k, l = pairs[cd]
            q, Q, hcd, Lcd = pdata[cd]
            pp, qq = p[:, None], q[None, :]
            alpha = pp * qq / (pp + qq)
            D = P[:, None, :] - Q[None, :, :]
            R = _hermite_R(Lab + Lcd, alpha, D[..., 0], D[..., 1], D[..., 2])
            pref = 2 * np.pi ** 2.5 / (pp * qq * np.sqrt(pp + qq))
            acc = 0.0
            for (t, u, v), eab in hab.items():
                for (tt, uu, vv), ecd in hcd.items():
                    sign = -1.0 if (tt + uu + vv) % 2 else 1.0
                    acc = acc + sign * eab[:, None] * ecd[None, :] * R[(t + tt, u + uu, v + vv)]
            val = float((pref * acc).sum())
            for (w, x) in ((i, j), (j, i)):
                for (y, z) in ((k, l), (l, k)):
                    g[w, x, y, z] = val
                    g[y, z, w, x] = val
    return g


# ------------------------------------------------------------------------------------------
class _HF:
    def __init__(self):
        self.mo_coeff = None
        self.mo_energy = None
        self.e_tot = None


class GtoMol:
    """Duck-typed ``Moldata_pyscf`` (``moldata_pyscf.py:19-62``) for an STO-3G molecule given as a
    Z-matrix or a list of ``(symbol, xyz Angstrom)``."""

    def __init__(self, geometry, basis="sto-3g"):
        if basis.lower() != "sto-3g":
            raise ValueError("only STO-3G parameters are tabulated here")
        atoms = from_zmatrix(geometry) if isinstance(geometry, str) else list(geometry)
        self.atoms = [(s, np.asarray(x, dtype=float) / BOHR) for s, x in atoms]      # Bohr
        self.charges = [STO3G[s]["Z"] for s, _ in self.atoms]
        self.nelectron = int(sum(self.charges))
        bas = Basis(self.atoms)
        n = bas.nao
        pos = [x for _, x in self.atoms]
        S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        for i in range(n):
            for j in range(i + 1):
                S[i, j] = S[j, i] = _overlap_pair(bas, i, j)
                T[i, j] = T[j, i] = _kinetic_pair(bas, i, j)
                V[i, j] = V[j, i] = _nuclear_pair(bas, i, j, self.charges, pos)
        self.overlap = S
        self.int1e_ao = T + V
        self.int2e_ao = _eri_all(bas)
        w, U = np.linalg.eigh(S)
        self.oao_coeff = U @ np.diag(w ** -0.5) @ U.T                                 # moldata_pyscf.py:13-16
        self.nuc = float(sum(self.charges[i] * self.charges[j] / np.linalg.norm(pos[i] - pos[j])
                             for i in range(len(pos)) for j in range(i)))
        self.nao = n
        self.hf = None

    def get_active_space_idx(self, ncas, nelecas):                                   # moldata_pyscf.py:42-56
        nelecore = self.nelectron - nelecas
        if nelecore % 2 == 1:
            raise ValueError('odd number of core electrons')
        occ_idx = np.arange(nelecore // 2)
        act_idx = (occ_idx[-1] + 1 + np.arange(ncas)) if len(occ_idx) > 0 else np.arange(ncas)
        virt_idx = np.arange(act_idx[-1] + 1, self.nao)
        return occ_idx, act_idx, virt_idx

    def run_rhf(self, verbose=0, conv_tol=1e-12, max_cycle=200):
        """Closed-shell SCF with DIIS from the core-Hamiltonian guess."""
        if self.hf is not None:
            return
        h, g, S, X = self.int1e_ao, self.int2e_ao, self.overlap, self.oao_coeff
        nocc = self.nelectron // 2

        def density(F):
            e, C = np.linalg.eigh(X.T @ F @ X)
            C = X @ C
            return 2 * C[:, :nocc] @ C[:, :nocc].T, e, C

        D, e, C = density(h)
        focks, errs, e_old = [], [], 0.0
        for it in range(max_cycle):
            J = np.einsum("pqrs,rs->pq", g, D)
            K = np.einsum("prqs,rs->pq", g, D)
            F = h + J - 0.5 * K
            e_tot = 0.5 * np.sum(D * (h + F)) + self.nuc
            err = X.T @ (F @ D @ S - S @ D @ F) @ X
            focks.append(F), errs.append(err)
            focks, errs = focks[-8:], errs[-8:]
            if len(focks) > 1:
                m = len(focks)
                Bm = -np.ones((m + 1, m + 1))
                Bm[m, m] = 0
                for a in range(m):
                    for b in range(m):
                        Bm[a, b] = np.sum(errs[a] * errs[b])
                rhs = np.zeros(m + 1)
                rhs[m] = -1
                try:
                    cvec = np.linalg.solve(Bm, rhs)[:m]
                    F = sum(ci * Fi for ci, Fi in zip(cvec, focks))
                except np.linalg.LinAlgError:
                    pass
            D, e, C = density(F)
            if verbose:
                print(f"scf {it:3d}  E = {e_tot:.12f}  |err| = {np.abs(err).max():.2e}")
            if abs(e_tot - e_old) < conv_tol and np.abs(err).max() < 1e-9:
                break
            e_old = e_tot
        # final orbitals from the converged (un-extrapolated) Fock matrix
        J = np.einsum("pqrs,rs->pq", g, D)
        K = np.einsum("prqs,rs->pq", g, D)
        F = h + J - 0.5 * K
        D, e, C = density(F)
        self.hf = _HF()
        self.hf.mo_coeff, self.hf.mo_energy = C, e
        self.hf.e_tot = float(0.5 * np.sum(D * (h + h + np.einsum("pqrs,rs->pq", g, D)
                                                - 0.5 * np.einsum("prqs,rs->pq", g, D))) + self.nuc)
