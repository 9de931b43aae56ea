"""TEST INFRASTRUCTURE ONLY -- CPU oracle (torch float64) of the auto_oo hot path.

This file is the checker, never the product: ``auto_oo_b200`` must not import
it.  It restates, stage by stage, the algorithm of the reference
(``/root/reference/src/auto_oo/oo_energy.py`` and
``utils/active_space.py:111-212``) with the *same torch CPU ops the reference
reaches through ``pennylane.math``* (``torch.einsum``, ``torch.linalg.matrix_exp``),
so that it is both the parity oracle and an honest CPU cost model of the
reference ("kind": "port" in bench.py).

Pinning: ``oracle/make_golden.py`` runs the VERBATIM reference (imported from
``/root/reference`` behind ``oracle/ref_shim.py``) and this file on the same
inputs and commits the reference outputs to ``tests/golden/*.npz``;
``tests/test_oracle.py`` re-checks this file against those fixtures and the
reference's own known-answer tests (``test/test_oo_energy.py:188-231``).
Molecular golden energies of the reference need PySCF integrals (absent here):
those are "parity unpinned" (DESIGN.md section 3).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import torch

DT = torch.float64


def _t(x):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=DT)


# --------------------------------------------------------------------------
# kappa packing                                  (oo_energy.py:63-118, 213-224)
# --------------------------------------------------------------------------
def tril_pairs(n):
    """Row-major strict-lower-triangle enumeration (l_t, r_t), l>r.
    oo_energy.py:82 / :93 (``np.tril_indices(size, k=-1)``)."""
    return np.tril_indices(n, k=-1)


def unpack_skew(vec):
    """vector -> skew matrix, K[l,r]=+v, K[r,l]=-v.  oo_energy.py:63-87."""
    vec = _t(vec)
    n = int(np.sqrt(8 * vec.shape[0] + 1) + 1) // 2
    rows, cols = tril_pairs(n)
    out = torch.zeros((n, n), dtype=DT)
    out[rows, cols] = vec
    out[cols, rows] = -vec
    return out


def pack_skew(mat):
    """skew matrix -> tril vector.  oo_energy.py:90-94."""
    rows, cols = tril_pairs(mat.shape[0])
    return mat[rows, cols]


def non_redundant_indices(occ_idx, act_idx, virt_idx, freeze_active):
    """Positions t in the tril enumeration whose pair is not occ-occ, not
    virt-virt and (if frozen) not act-act.  oo_energy.py:97-118."""
    occ, act, virt = (set(int(i) for i in x) for x in (occ_idx, act_idx, virt_idx))
    nao = len(occ) + len(act) + len(virt)
    keep = []
    for t, (l, r) in enumerate(zip(*tril_pairs(nao))):
        l, r = int(l), int(r)
        if l in occ and r in occ:
            continue
        if l in virt and r in virt:
            continue
        if freeze_active and l in act and r in act:
            continue
        keep.append(t)
    no, na, nv = len(occ), len(act), len(virt)
    expect = no * na + na * nv + no * nv + (0 if freeze_active else na * (na - 1) // 2)
    assert expect == len(keep)                                   # oo_energy.py:117
    return np.asarray(keep, dtype=int)


def active_space_idx(nao, nelec, ncas, nelecas):
    """occ / act / virt index ranges.  moldata_pyscf.py:42-56."""
    ncore_e = nelec - nelecas
    if ncore_e % 2 == 1:
        raise ValueError('odd number of core electrons')         # :47-48
    occ = np.arange(ncore_e // 2)
    act = (occ[-1] + 1 + np.arange(ncas)) if len(occ) > 0 else np.arange(ncas)
    virt = np.arange(act[-1] + 1, nao)
    return occ, act, virt


def kappa_to_skew(kappa, params_idx, nao):
    """oo_energy.py:213-219."""
    full = torch.zeros(nao * (nao - 1) // 2, dtype=DT)
    full[torch.as_tensor(np.asarray(params_idx), dtype=torch.long)] = _t(kappa)
    return unpack_skew(full)



def kappa_to_skew_diff(kappa, params_idx, nao):
    """Same map as :func:`kappa_to_skew`, written out-of-place so autograd can differentiate
    through it (used only by the analytic-vs-autograd tests)."""
    rows, cols = tril_pairs(nao)
    pidx = np.asarray(params_idx, dtype=int)
    L = torch.as_tensor(rows[pidx], dtype=torch.long)
    R = torch.as_tensor(cols[pidx], dtype=torch.long)
    out = torch.zeros((nao, nao), dtype=DT)
    out = out.index_put((L, R), kappa)
    return out.index_put((R, L), -kappa)

def skew_to_kappa(mat, params_idx):
    """oo_energy.py:221-224."""
    return pack_skew(mat)[torch.as_tensor(np.asarray(params_idx), dtype=torch.long)]


def rotation_from_kappa(kappa, params_idx, nao):
    """U = expm(-K).  oo_energy.py:226-230 (math.expm -> torch.linalg.matrix_exp)."""
    return torch.linalg.matrix_exp(-kappa_to_skew(kappa, params_idx, nao))


# --------------------------------------------------------------------------
# integral transforms                                     (oo_energy.py:21-51)
# --------------------------------------------------------------------------
def transform_1e(h_ao, C):
    """C^T h C.  oo_energy.py:44-46."""
    h_ao, C = _t(h_ao), _t(C)
    return C.T @ h_ao @ C


def transform_4index(M, C0, C1, C2, C3):
    """Four sequential quarter transforms with implicit-output einsums
    (alphabetical output order, i.e. the result stays [i,j,k,l]).
    oo_energy.py:21-30."""
    M, C0, C1, C2, C3 = (_t(x) for x in (M, C0, C1, C2, C3))
    M = torch.einsum('pi,pqrs->iqrs', C0, M)
    M = torch.einsum('qj,iqrs->ijrs', C1, M)
    M = torch.einsum('rk,ijrs->ijks', C2, M)
    M = torch.einsum('sl,ijks->ijkl', C3, M)
    return M


def transform_2e(g_ao, C):
    """oo_energy.py:33-41, :49-51."""
    return transform_4index(g_ao, C, C, C, C)


def mo_ao_to_mo_oao(mo_coeff, overlap):
    """S^{1/2} C via eigh(S) (numpy).  oo_energy.py:54-60."""
    w, v = np.linalg.eigh(np.asarray(overlap))
    return (v @ np.diag(w ** 0.5) @ v.T) @ np.asarray(mo_coeff)


# --------------------------------------------------------------------------
# active-space Hamiltonian + energy     (active_space.py:111-212, oo_energy.py:178-211)
# --------------------------------------------------------------------------
def active_space_integrals(h, g, occ_idx, act_idx):
    """(E_core, h~, g_act).  active_space.py:111-174."""
    occ = np.asarray(occ_idx, dtype=int)
    act = np.asarray(act_idx, dtype=int)
    oa = np.ix_(act, act)
    ta = np.ix_(act, act, act, act)
    core = (2 * torch.sum(h[occ, occ])                                   # :148
            + 2 * torch.sum(g[occ, occ, :, :][:, occ, occ])              # :149-151
            - torch.sum(g[occ, :, :, occ][:, occ, occ]))                 # :152-154
    g_act = g[ta]                                                        # :158
    h_act = (h[oa]                                                       # :163
             + 2 * torch.sum(g[:, :, occ, occ][act, :, :][:, act, :], dim=2)   # :164-166
             - torch.sum(g[:, occ, occ, :][act, :, :][:, :, act], dim=1))      # :167-168
    return core, h_act, g_act


def hamiltonian_coefficients(e_nuc, h, g, occ_idx=None, act_idx=None):
    """(c0, c1, c2) = (E_core + E_nuc, h~, g_act/2).  active_space.py:177-212."""
    if occ_idx is None and act_idx is None:
        c0 = e_nuc
    else:
        core, h, g = active_space_integrals(h, g, occ_idx, act_idx)
        c0 = core + e_nuc
    return c0, h, 0.5 * g


def energy_from_coefficients(c0, c1, c2, one_rdm, two_rdm):
    """E = c0 + <c1,gamma> + <c2,Gamma>.  oo_energy.py:194-197."""
    return (c0 + torch.einsum('pq,pq->', c1, _t(one_rdm))
            + torch.einsum('pqrs,pqrs->', c2, _t(two_rdm)))


# --------------------------------------------------------------------------
# Fock matrices and gradient                           (oo_energy.py:238-309)
# --------------------------------------------------------------------------
def fock_core(h, g, occ_idx):
    """F^I_mn = h_mn + sum_i (2 g_mnii - g_miin).  oo_energy.py:272-284."""
    occ = np.asarray(occ_idx, dtype=int)
    return h + (2 * torch.sum(g[:, :, occ, occ], dim=2)
                - torch.sum(g[:, occ, occ, :], dim=1))


def fock_active(g, one_rdm, act_idx):
    """F^A_mn = sum_vw gamma_vw (g_mnvw - g_mwvn / 2).  oo_energy.py:286-298."""
    act = np.asarray(act_idx, dtype=int)
    gt = (g[:, :, :, act][:, :, act, :]
          - 0.5 * g[:, :, act, :][:, act, :, :].permute(0, 3, 2, 1))
    return torch.einsum('vw,mnvw->mn', _t(one_rdm), gt)


def fock_generalized(h, g, one_rdm, two_rdm, occ_idx, act_idx):
    """Generalized Fock matrix.  oo_energy.py:238-270."""
    occ = np.asarray(occ_idx, dtype=int)
    act = np.asarray(act_idx, dtype=int)
    one_rdm, two_rdm = _t(one_rdm), _t(two_rdm)
    f_i = fock_core(h, g, occ)
    f_a = fock_active(g, one_rdm, act)
    out = torch.zeros_like(h)
    out[occ] = 2 * (f_i[:, occ] + f_a[:, occ]).T                          # :262-264
    g_nact = g[:, :, :, act][:, :, act, :][:, act, :, :]                  # :270
    out[act] = (torch.einsum('nw,vw->vn', f_i[:, act], one_rdm)           # :265-267
                + torch.einsum('vwxy,nwxy->vn', two_rdm, g_nact))
    return out


def gradient_matrix(h, g, one_rdm, two_rdm, occ_idx, act_idx):
    """G = 2 (F - F^T).  oo_energy.py:300-309."""
    f = fock_generalized(h, g, one_rdm, two_rdm, occ_idx, act_idx)
    return 2 * (f - f.T)


# --------------------------------------------------------------------------
# Hessian, reference formulation (dense N^4 / N^6)     (oo_energy.py:311-402)
# --------------------------------------------------------------------------
def full_rdms(one_rdm, two_rdm, nao, occ_idx, act_idx):
    """Embed active RDMs into the full orbital space.  oo_energy.py:342-379."""
    occ = np.asarray(occ_idx, dtype=int)
    act = np.asarray(act_idx, dtype=int)
    one_rdm, two_rdm = _t(one_rdm), _t(two_rdm)
    no = len(occ)
    eye = torch.eye(no, dtype=DT)
    d1 = torch.zeros((nao, nao), dtype=DT)
    d2 = torch.zeros((nao,) * 4, dtype=DT)
    d1[occ, occ] = 2.0                                                    # :359-360
    d1[np.ix_(act, act)] = one_rdm                                        # :361
    d2[np.ix_(occ, occ, occ, occ)] = (4 * torch.einsum('ij,kl->ijkl', eye, eye)
                                      - 2 * torch.einsum('il,jk->ijkl', eye, eye))  # :363-365
    d2[np.ix_(occ, occ, act, act)] = 2 * torch.einsum('wv,ij->ijwv', one_rdm, eye)  # :366-368
    d2[np.ix_(act, act, occ, occ)] = 2 * torch.einsum('wv,ij->wvij', one_rdm, eye)  # :369-371
    d2[np.ix_(occ, act, act, occ)] = -torch.einsum('wv,ij->iwvj', one_rdm, eye)     # :372-374
    d2[np.ix_(act, occ, occ, act)] = -torch.einsum('wv,ij->vjiw', one_rdm, eye)     # :375-377
    d2[np.ix_(act, act, act, act)] = two_rdm                              # :378
    return d1, d2


def y_matrix(g, d2_full):
    """Y_pqrs = sum_mn[(G_pmrn + G_pmnr) g_qmns + G_prmn g_qsmn].  oo_energy.py:381-393."""
    return (torch.einsum('pmrn,qmns->pqrs', d2_full, g)
            + torch.einsum('pmnr,qmns->pqrs', d2_full, g)
            + torch.einsum('prmn,qsmn->pqrs', d2_full, g))


def hessian_full(h, g, one_rdm, two_rdm, occ_idx, act_idx):
    """Rank-4 orbital Hessian.  oo_energy.py:311-340."""
    nao = h.shape[0]
    d1, d2 = full_rdms(one_rdm, two_rdm, nao, occ_idx, act_idx)
    y = y_matrix(g, d2)
    f = fock_generalized(h, g, one_rdm, two_rdm, occ_idx, act_idx)
    fs = f + f.T
    x = (2 * torch.einsum('pr,qs->pqrs', d1, h)
         - torch.einsum('pr,qs->pqrs', fs, torch.eye(nao, dtype=DT))
         + 2 * y)
    return x - x.permute(0, 1, 3, 2) - x.permute(1, 0, 2, 3) + x.permute(1, 0, 3, 2)


def hessian_to_matrix(full_hess, params_idx):
    """(N,N,N,N) -> (n_kappa, n_kappa).  oo_energy.py:395-402."""
    nao = full_hess.shape[0]
    rows, cols = tril_pairs(nao)
    part = full_hess[rows, cols, :, :][:, rows, cols]
    pidx = np.asarray(params_idx, dtype=int)
    return part[pidx, :][:, pidx]


# --------------------------------------------------------------------------
# Hessian, I-space evaluation -- same numbers as
# hessian_to_matrix(hessian_full(...)) but O(nI^4 N^2); used as the oracle at
# sizes where the dense N^6 form cannot run.  Checked against the dense form in
# tests/test_oracle.py.
# --------------------------------------------------------------------------
def hessian_matrix_ispace(h, g, one_rdm, two_rdm, occ_idx, act_idx, params_idx):
    nao = h.shape[0]
    occ = np.asarray(occ_idx, dtype=int)
    act = np.asarray(act_idx, dtype=int)
    ni = len(occ) + len(act)
    assert np.array_equal(np.concatenate([occ, act]), np.arange(ni)), \
        "I-space form assumes occ|act are the leading contiguous indices"
    d1, d2 = full_rdms(one_rdm, two_rdm, ni, occ, act)       # exact: zero outside I
    a1 = d2.permute(0, 2, 1, 3) + d2.permute(0, 3, 1, 2)     # [(p r),(m n)] = G_pmrn + G_pmnr
    a2 = d2                                                  # [(p r),(m n)] = G_prmn
    kc = g[:, :ni, :ni, :].permute(1, 2, 0, 3)               # [(m n),(q s)] = g_qmns
    jc = g[:, :, :ni, :ni].permute(2, 3, 0, 1)               # [(m n),(q s)] = g_qsmn
    y = (a1.reshape(ni * ni, ni * ni) @ kc.reshape(ni * ni, nao * nao)
         + a2.reshape(ni * ni, ni * ni) @ jc.reshape(ni * ni, nao * nao))
    t = 2 * y.reshape(ni, ni, nao, nao) + 2 * torch.einsum('pr,qs->prqs', d1, h)
    f = fock_generalized(h, g, one_rdm, two_rdm, occ, act)
    fs = f + f.T

    rows, cols = tril_pairs(nao)
    pidx = np.asarray(params_idx, dtype=int)
    L = torch.as_tensor(rows[pidx], dtype=torch.long)
    R = torch.as_tensor(cols[pidx], dtype=torch.long)

    def x(p, q, r, s):
        """X(p,q,r,s) on outer-product index grids (n_kappa x n_kappa)."""
        P, Rr = p[:, None], r[None, :]
        Q, S = q[:, None], s[None, :]
        out = -fs[P, Rr] * (Q == S).to(DT)
        inside = (P < ni) & (Rr < ni)
        tt = t[torch.clamp(P, max=ni - 1), torch.clamp(Rr, max=ni - 1), Q, S]
        return out + torch.where(inside, tt, torch.zeros_like(tt))

    return x(L, R, L, R) - x(L, R, R, L) - x(R, L, L, R) + x(R, L, R, L)


# --------------------------------------------------------------------------
# A problem instance = what OO_energy holds          (oo_energy.py:121-236)
# --------------------------------------------------------------------------
class OracleProblem:
    """Inputs of one evaluation and the reference's three entry points
    (energy_from_kappa, analytic_gradient -> vector, analytic_hessian -> matrix)."""

    def __init__(self, int1e_ao, int2e_ao, oao_coeff, oao_mo_coeff, nuc, nelec,
                 ncas, nelecas, freeze_active=False):
        self.h_ao, self.g_ao = _t(int1e_ao), _t(int2e_ao)
        self.oao_coeff, self.oao_mo_coeff = _t(oao_coeff), _t(oao_mo_coeff)
        self.nuc = float(nuc)
        self.nao = self.h_ao.shape[0]
        self.occ_idx, self.act_idx, self.virt_idx = active_space_idx(
            self.nao, nelec, ncas, nelecas)
        self.params_idx = non_redundant_indices(
            self.occ_idx, self.act_idx, self.virt_idx, freeze_active)
        self.n_kappa = len(self.params_idx)

    @property
    def mo_coeff(self):                                              # oo_energy.py:173-176
        return self.oao_coeff @ self.oao_mo_coeff

    def rotated_mo(self, kappa=None):                                # :199-201, :232-236
        if kappa is None:
            return self.mo_coeff
        return self.mo_coeff @ rotation_from_kappa(kappa, self.params_idx, self.nao)

    def mo_integrals(self, kappa=None):
        C = self.rotated_mo(kappa)
        return transform_1e(self.h_ao, C), transform_2e(self.g_ao, C)

    def active_integrals(self, kappa=None):                          # :204-211
        h, g = self.mo_integrals(kappa)
        return hamiltonian_coefficients(self.nuc, h, g, self.occ_idx, self.act_idx)

    def energy(self, one_rdm, two_rdm, kappa=None):                  # :178-202
        c0, c1, c2 = self.active_integrals(kappa)
        return energy_from_coefficients(c0, c1, c2, one_rdm, two_rdm)

    def gradient(self, one_rdm, two_rdm, kappa=None):                # :404-413 + :221-224
        h, g = self.mo_integrals(kappa)
        G = gradient_matrix(h, g, one_rdm, two_rdm, self.occ_idx, self.act_idx)
        return skew_to_kappa(G, self.params_idx)

    def hessian(self, one_rdm, two_rdm, kappa=None, ispace=False):   # :415-424 + :395-402
        h, g = self.mo_integrals(kappa)
        if ispace:
            return hessian_matrix_ispace(h, g, one_rdm, two_rdm, self.occ_idx,
                                         self.act_idx, self.params_idx)
        return hessian_to_matrix(
            hessian_full(h, g, one_rdm, two_rdm, self.occ_idx, self.act_idx),
            self.params_idx)

    def evaluate(self, one_rdm, two_rdm, kappa=None, ispace=False):
        """(E, G, H) the way the reference computes them: three independent
        passes, each redoing the 4-index transform (oo_energy.py:207-208,
        :410-411, :421-422)."""
        return (self.energy(one_rdm, two_rdm, kappa),
                self.gradient(one_rdm, two_rdm, kappa),
                self.hessian(one_rdm, two_rdm, kappa, ispace=ispace))
