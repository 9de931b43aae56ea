"""TEST INFRASTRUCTURE ONLY -- CPU oracle of E, gradient and Hessian that scales to the headline size.

``oracle/oo_oracle.py`` restates the reference literally (complete four-index transform, dense N^6
Y-matrix); that form needs a dozen N^4 tensors and cannot run at N = 256.  This
file computes THE SAME NUMBERS from the two integral classes the formulas actually read,

    Jc[m,n,a,b] = g'[a,b,m,n] = (ab|mn)        Kc[n,m,a,b] = g'[a,m,n,b] = (am|nb)

with m, n in I = occ + act and a, b general, plus the I-space Hessian of SURVEY Appendix A.6.  Every
function cites the reference lines whose numbers it reproduces.  It is written with plain torch CPU
matmuls / einsums in its own association order (it shares no code with ``auto_oo_b200``).

Pinning: ``tests/test_oracle.py`` checks it against the verbatim-reference fixtures
(``tests/golden/*.npz``, N = 7 ... 43) and ``oracle/make_golden_large.py`` checks it against the
verbatim reference at N = 114 before it is trusted at N = 256.  Precondition (asserted): the AO
integrals have the 8-fold symmetry of real orbitals -- the class extraction uses
(am|nb) = (ma|bn) -- which is also the precondition under which the reference's analytic formulas are
derivatives of its energy.
"""
from __future__ import annotations

import numpy as np
import torch

from . import oo_oracle as orc

DT = torch.float64


def _t(x):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=DT)


def class_integrals(h_ao, g_ao, C, ni, check_symmetry=True):
    """(h', Jc, Kc) at MO coefficients ``C``: the slices of ``general_4index_transform`` /
    ``int1e_transform`` (oo_energy.py:21-51) that energy, gradient and Hessian read.
    Peak memory: g_ao + ni N^3 doubles."""
    h_ao, g_ao, C = _t(h_ao), _t(g_ao), _t(C)
    N = C.shape[0]
    if check_symmetry:
        probe = g_ao[: min(N, 6)]                     # slices only: no N^4 temporary
        tol = 1e-12 * float(probe.abs().max())
        assert torch.allclose(probe, probe.transpose(2, 3), rtol=0, atol=tol), "class_oracle needs (pq|rs) = (pq|sr)"
        assert torch.allclose(g_ao[:4, :4], g_ao[:, :, :4, :4].permute(2, 3, 0, 1), rtol=0, atol=tol), \
            "class_oracle needs (pq|rs) = (rs|pq)"
    CI = C[:, :ni]
    # T1[m,q,r,s] = sum_p C[p,m] g[p,q,r,s]                                (oo_energy.py:26)
    T1 = (CI.T @ g_ao.reshape(N, N * N * N)).reshape(ni, N, N, N)
    # Coulomb class: (mn|ab), then the pair swap (ab|mn)                        (:27-29)
    T2 = torch.einsum('qn,mqrs->mnrs', CI, T1)
    J = torch.matmul(torch.matmul(C.T, T2), C)                            # [m,n,a,b] = sum_rs C_ra C_sb T2[m,n,r,s]
    del T2
    # exchange class: (ma|bn) = (am|nb)
    T3 = torch.matmul(T1, CI)                                             # [m,q,r,n]
    del T1
    T3 = torch.einsum('qa,mqrn->marn', C, T3)
    Kx = torch.einsum('rb,marn->mabn', C, T3)                             # g'[m,a,b,n]
    del T3
    K = Kx.permute(3, 0, 1, 2).contiguous()                               # Kc[n,m,a,b] = g'[a,m,n,b]
    return C.T @ h_ao @ C, J.contiguous(), K


class ClassEvaluation:
    """E, generalized Fock, gradient and Hessian at one set of MO coefficients from (h', Jc, Kc)."""

    def __init__(self, h, J, K, nuc, no, na, params_idx):
        self.h, self.J, self.K = h, J, K
        self.nuc, self.no, self.na = float(nuc), int(no), int(na)
        self.ni = self.no + self.na
        self.N = h.shape[0]
        self.params_idx = np.asarray(params_idx, dtype=int)
        self.act = slice(self.no, self.ni)

    # ---- active-space Hamiltonian                      (utils/active_space.py:147-174, :209-212)
    def hamiltonian(self):
        no, act = self.no, self.act
        io = torch.arange(no)
        Jd = self.J[io, io]                                   # [i, a, b] = g'[a,b,i,i]
        Kd = self.K[io, io]                                   # [i, a, b] = g'[a,i,i,b]
        c0 = (self.nuc + 2 * self.h[io, io].sum() + 2 * Jd[:, io, io].sum() - Kd[:, io, io].sum())
        c1 = self.h[act, act] + 2 * Jd[:, act, act].sum(0) - Kd[:, act, act].sum(0)
        c2 = 0.5 * self.J[act, act][:, :, act, act].permute(2, 3, 0, 1)   # g'[t,u,v,w] = Jc[v,w,t,u]
        return c0, c1, c2

    def energy(self, one, two):                               # oo_energy.py:194-197
        c0, c1, c2 = self.hamiltonian()
        return c0 + (c1 * _t(one)).sum() + (c2 * _t(two)).sum()

    # ---- Fock matrices                                                   (oo_energy.py:238-298)
    def fock_core(self):
        io = torch.arange(self.no)
        return self.h + 2 * self.J[io, io].sum(0) - self.K[io, io].sum(0)

    def fock_active(self, one):
        act = self.act
        one = _t(one)
        # g_mnvw = Jc[v,w,m,n];  g_mwvn = g'[m,w,v,n] = Kc[v,w,m,n]
        return (torch.einsum('vw,vwmn->mn', one, self.J[act, act])
                - 0.5 * torch.einsum('vw,vwmn->mn', one, self.K[act, act]))

    def fock_generalized(self, one, two):
        no, act = self.no, self.act
        one, two = _t(one), _t(two)
        fi, fa = self.fock_core(), self.fock_active(one)
        F = torch.zeros_like(self.h)
        F[:no] = 2 * (fi[:, :no] + fa[:, :no]).T                              # :262-264
        g_nwxy = self.J[act, act][:, :, :, act]                              # [x,y,n,w] = g'[n,w,x,y]
        F[act] = (torch.einsum('nw,vw->vn', fi[:, act], one)                  # :265-270
                  + torch.einsum('vwxy,xynw->vn', two, g_nwxy))
        return F

    def gradient(self, one, two):                             # :300-309 + :221-224
        F = self.fock_generalized(one, two)
        return orc.skew_to_kappa(2 * (F - F.T), self.params_idx)

    # ---- Hessian                                           (oo_energy.py:311-402; Appendix A.6)
    def _t_matrix(self, one, two):
        ni, N = self.ni, self.N
        occ, actl = np.arange(self.no), np.arange(self.no, ni)
        d1, d2 = orc.full_rdms(one, two, ni, occ, actl)       # exact: the full-space RDMs vanish outside I
        a1 = (d2.permute(0, 2, 1, 3) + d2.permute(0, 3, 1, 2)).reshape(ni * ni, ni * ni)   # G_pmrn + G_pmnr
        a2 = d2.reshape(ni * ni, ni * ni)                                                  # G_prmn
        kc = self.K.permute(1, 0, 2, 3).reshape(ni * ni, N * N)   # [(m n),(q s)] = g'[q,m,n,s] = Kc[n,m,q,s]
        jc = self.J.reshape(ni * ni, N * N)                       # [(m n),(q s)] = g'[q,s,m,n]
        y = a1 @ kc + a2 @ jc
        return 2 * y.reshape(ni, ni, N, N) + 2 * torch.einsum('pr,qs->prqs', d1, self.h)

    def hessian(self, one, two, rows=None):
        """(n_kappa, n_kappa) Hessian matrix (or the given rows of it)."""
        ni, N = self.ni, self.N
        t = self._t_matrix(one, two)
        F = self.fock_generalized(one, two)
        fs = F + F.T
        rr, cc = orc.tril_pairs(N)
        L = torch.as_tensor(rr[self.params_idx], dtype=torch.long)
        R = torch.as_tensor(cc[self.params_idx], dtype=torch.long)
        Lr, Rr_ = (L, R) if rows is None else (L[rows], R[rows])

        def x(p, q, r, s):
            P, Rr = p[:, None], r[None, :]
            Q, S = q[:, None], s[None, :]
            out = -fs[P, Rr] * (Q == S).to(DT)
            inside = (P < ni) & (Rr < ni)
            tt = t[torch.clamp(P, max=ni - 1), torch.clamp(Rr, max=ni - 1), Q, S]
            return out + torch.where(inside, tt, torch.zeros_like(tt))

        return x(Lr, Rr_, L, R) - x(Lr, Rr_, R, L) - x(Rr_, Lr, L, R) + x(Rr_, Lr, R, L)


class ClassProblem:
    """Same constructor and entry points as :class:`oracle.oo_oracle.OracleProblem`."""

    def __init__(self, int1e_ao, int2e_ao, oao_coeff, oao_mo_coeff, nuc, nelec, ncas, nelecas,
                 freeze_active=False):
        self.h_ao, self.g_ao = _t(int1e_ao), _t(int2e_ao)
        self.oao_coeff, self.oao_mo_coeff = _t(oao_coeff), _t(oao_mo_coeff)
        self.nuc = float(nuc)
        self.nao = self.h_ao.shape[0]
        self.occ_idx, self.act_idx, self.virt_idx = orc.active_space_idx(self.nao, nelec, ncas, nelecas)
        self.params_idx = orc.non_redundant_indices(self.occ_idx, self.act_idx, self.virt_idx, freeze_active)
        self.n_kappa = len(self.params_idx)
        self.no, self.na = len(self.occ_idx), len(self.act_idx)

    def at(self, kappa=None):
        C = self.oao_coeff @ self.oao_mo_coeff                               # oo_energy.py:173-176
        if kappa is not None:
            C = C @ orc.rotation_from_kappa(kappa, self.params_idx, self.nao)  # :199-201, :226-236
        h, J, K = class_integrals(self.h_ao, self.g_ao, C, self.no + self.na)
        return ClassEvaluation(h, J, K, self.nuc, self.no, self.na, self.params_idx)

    def evaluate(self, one_rdm, two_rdm, kappa=None):
        ev = self.at(kappa)
        return ev.energy(one_rdm, two_rdm), ev.gradient(one_rdm, two_rdm), ev.hessian(one_rdm, two_rdm)
