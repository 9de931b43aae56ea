"""Synthetic molecular data of the BASELINE config shapes and PennyLane-free
RDM providers.

The reference obtains its inputs from PySCF (``moldata_pyscf.py:19-56``) and
from a PennyLane state vector (``pqc.py:192-222``); neither is in scope (nor
installed).  This module supplies objects with the same duck-typed interface:

* :class:`SyntheticMol` -- attributes ``int1e_ao, int2e_ao, overlap, oao_coeff,
  nuc, nao`` and ``get_active_space_idx`` exactly as ``OO_energy.__init__``
  reads them (reference ``oo_energy.py:154-165``); integrals have the 8-fold
  symmetry the analytic gradient/Hessian assume.
* :func:`random_rdms` -- symmetric random 1-/2-RDMs of the required index
  symmetries; :class:`CIVectorCircuit` -- N-representable spin-summed RDMs
  ``<E_pq>``, ``<e_pqrs>`` (``utils/active_space.py:29-83``) of a parameterised
  CI vector, differentiable in ``theta`` (stand-in for ``Parameterized_circuit``:
  ``get_rdms(theta)``, ``theta_shape``).
"""
from __future__ import annotations

import itertools
import math as _math

import numpy as np
import torch

# (name, nao, nelec, ncas, nelecas) -- BASELINE.json configs / SURVEY section 8 table
CONFIG_SHAPES = {
    "h2o_sto3g_cas44": (7, 10, 4, 4),
    "ch2nh_631gs_cas44": (34, 16, 4, 4),
    "n2_ccpvdz_cas66": (28, 14, 6, 6),
    "c6h6_ccpvdz_cas66": (114, 42, 6, 6),
    "synthetic_n256_cas1212": (256, 76, 12, 12),
    # beyond BASELINE: capacity shapes for the pair-sharded evaluation (bench.py --mode strong)
    "synthetic_n384_cas1212": (384, 76, 12, 12),
    "synthetic_n512_cas1212": (512, 76, 12, 12),
}


def _sym(a):
    return 0.5 * (a + a.transpose(-1, -2))


class SyntheticMol:
    """Seeded synthetic stand-in for ``Moldata_pyscf``.

    ``h = sym(randn)/2 - diag(linspace(10, .5, N))``;
    ``g = 0.3 * B B^T`` with ``B[pq,P]`` symmetric in ``pq`` (density-fitting
    shape, rank ``2N``) so ``g`` is 8-fold symmetric and positive;
    ``S = I + 0.1 sym(randn)/sqrt(N)``, ``oao_coeff = S^{-1/2}``; ``nuc = 9``.
    Tensors are float64 on ``device``; numpy views are exposed for CPU.
    ``rng_device``: where the random numbers are drawn (default: ``device``).  ``rng_device="cpu"`` with a
    CUDA ``device`` gives the same ``h, S, B, C_oao`` as the all-CPU object of the same seed (the golden
    fixtures are produced on the CPU); only the product ``B B^T`` is then formed on the device.
    """

    def __init__(self, nao, nelec, seed=0, device="cpu", eri_rank=None, build_eri=True, rng_device=None):
        self.nao = int(nao)
        self.nelectron = int(nelec)
        self.seed = int(seed)
        out_dev = torch.device(device)
        dev = out_dev if rng_device is None else torch.device(rng_device)
        gen = torch.Generator(device=dev).manual_seed(20240 + self.seed)
        N = self.nao
        R = int(eri_rank) if eri_rank else 2 * N
        kw = dict(dtype=torch.float64, device=dev, generator=gen)

        h = _sym(torch.randn(N, N, **kw)) * 0.5
        h = h - torch.diag(torch.linspace(10.0, 0.5, N, dtype=torch.float64, device=dev))
        S = torch.eye(N, dtype=torch.float64, device=dev) + 0.1 * _sym(torch.randn(N, N, **kw)) / _math.sqrt(N)
        w, v = torch.linalg.eigh(S.cpu())            # host eigh: identical on any device
        X = ((v * w.pow(-0.5)) @ v.T).to(dev)
        B = torch.randn(N, N, R, **kw)
        B = 0.5 * (B + B.transpose(0, 1)) / _math.sqrt(R)
        c = torch.randn(N, N, **kw)                  # drawn here so the stream does not depend on build_eri
        self._B = B.to(out_dev)
        self._int1e = h.to(out_dev)
        self._overlap = S.to(out_dev)
        self._oao = X.to(out_dev)
        self.nuc = 9.0
        self._int2e = None
        if build_eri:
            self._int2e = self.build_eri()
        q, r = torch.linalg.qr(c.cpu())
        q = q * torch.sign(torch.diagonal(r))[None, :]
        self._oao_mo = q.to(out_dev)

    def build_eri(self, out=None):
        """g[p,q,r,s] = 0.3 * sum_P B[p,q,P] B[r,s,P]; chunked so N=256 builds on device."""
        N = self.nao
        B2 = self._B.reshape(N * N, -1)
        g = out if out is not None else torch.empty(N * N, N * N, dtype=torch.float64, device=B2.device)
        g = g.view(N * N, N * N)
        step = max(1, (1 << 27) // (N * N))
        for a in range(0, N * N, step):
            torch.matmul(B2[a:a + step], B2.T, out=g[a:a + step])
        g.mul_(0.3)
        return g.view(N, N, N, N)

    def int2e_packed_slab(self, lo, hi):
        """Columns ``[lo, hi)`` of the 8-fold packed integrals ``g8[RS][PQ] = g[r,s,p,q]`` (``r >= s``, ``p >= q``,
        pairs over the orbitals padded to even, zero padding) straight from the density-fitting factor -- the N^4
        tensor is never formed.  What ``OO_energy(..., shard="pairs")`` asks every rank for
        (:class:`auto_oo_b200.distributed.PairShard`)."""
        N = self.nao
        ld = N + (N & 1)
        B = self._B
        if ld > N:
            Bp = torch.zeros(ld, ld, B.shape[2], dtype=B.dtype, device=B.device)
            Bp[:N, :N] = B
            B = Bp
        rows, cols = np.tril_indices(ld)
        rows_t = torch.as_tensor(rows, device=B.device)
        cols_t = torch.as_tensor(cols, device=B.device)
        packed = B[rows_t, cols_t]                                  # (npair, R)
        slab = torch.zeros(hi - lo, packed.shape[1], dtype=B.dtype, device=B.device)
        top = min(hi, packed.shape[0])
        if top > lo:
            slab[: top - lo] = packed[lo:top]
        return (packed @ slab.T).mul_(0.3)

    # -- duck-typed Moldata_pyscf attributes (numpy on CPU, like PySCF arrays) --
    def _np(self, t):
        return t.cpu().numpy() if t.device.type == "cpu" else t

    @property
    def int1e_ao(self):
        return self._np(self._int1e)

    @property
    def int2e_ao(self):
        return self._np(self._int2e)

    @property
    def overlap(self):
        return self._np(self._overlap)

    @property
    def oao_coeff(self):
        return self._np(self._oao)

    @property
    def random_oao_mo_coeff(self):
        """A seeded orthogonal OAO->MO matrix (stands in for the RHF default)."""
        return self._np(self._oao_mo)

    def get_active_space_idx(self, ncas, nelecas):
        """Same rule as ``moldata_pyscf.py:42-56``."""
        nelecore = self.nelectron - nelecas
        if nelecore % 2 == 1:
            raise ValueError('odd number of core electrons')
        occ_idx = np.arange(nelecore // 2)
        act_idx = (occ_idx[-1] + 1 + np.arange(ncas)) if len(occ_idx) > 0 else np.arange(ncas)
        virt_idx = np.arange(act_idx[-1] + 1, self.nao)
        return occ_idx, act_idx, virt_idx


def random_rdms(ncas, nelecas, seed=0, device="cpu"):
    """gamma = gamma^T, Gamma_pqrs = Gamma_rspq = Gamma_qpsr (the symmetries the
    analytic formulas assume); magnitudes of a typical CAS."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(7700 + int(seed))
    kw = dict(dtype=torch.float64, device=dev, generator=gen)
    one = _sym(torch.randn(ncas, ncas, **kw)) * 0.1
    one = one + torch.eye(ncas, dtype=torch.float64, device=dev) * (nelecas / ncas)
    two = torch.randn(ncas, ncas, ncas, ncas, **kw)
    two = two + two.permute(2, 3, 0, 1)
    two = two + two.permute(1, 0, 3, 2)
    return one, two * 0.025


def random_kappa(n_kappa, seed=0, scale=0.05, device="cpu", batch=None):
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(9100 + int(seed))
    shape = (n_kappa,) if batch is None else (batch, n_kappa)
    return torch.randn(*shape, dtype=torch.float64, device=dev, generator=gen) * scale


# --------------------------------------------------------------------------
# CI-vector RDM provider
# --------------------------------------------------------------------------
def _strings(norb, nocc):
    return [sum(1 << i for i in c) for c in itertools.combinations(range(norb), nocc)]


def _excitation_tables(norb, nocc):
    """Dense (norb, norb, D, D) matrices of a+_p a_q on the string space."""
    strs = _strings(norb, nocc)
    pos = {s: i for i, s in enumerate(strs)}
    D = len(strs)
    E = np.zeros((norb, norb, D, D))
    for j, s in enumerate(strs):
        for q in range(norb):
            if not (s >> q) & 1:
                continue
            s1 = s & ~(1 << q)
            sign_q = (-1) ** bin(s & ((1 << q) - 1)).count("1")
            for p in range(norb):
                if (s1 >> p) & 1:
                    continue
                sign_p = (-1) ** bin(s1 & ((1 << p) - 1)).count("1")
                E[p, q, pos[s1 | (1 << p)], j] = sign_q * sign_p
    return E


class CIVectorCircuit:
    """theta -> normalised CI vector ``expm(sum_k theta_k A_k) |HF>`` (``A_k`` fixed
    seeded antisymmetric generators) -> spin-summed RDMs.

    Mirrors the two members ``OO_pqc`` uses from ``Parameterized_circuit``
    (``oo_pqc.py:83``, ``:90-93``): ``get_rdms(theta)`` and ``theta_shape``.
    gamma_pq = <E_pq>, Gamma_pqrs = <E_pq E_rs> - delta_qr <E_ps>
    (``utils/active_space.py:57-83``).
    """

    def __init__(self, ncas, nelecas, n_theta=2, seed=0, interface="torch"):
        if isinstance(nelecas, (tuple, list)):
            na, nb = nelecas
        else:
            na, nb = (nelecas + 1) // 2, nelecas // 2
        self.ncas = ncas
        self.nelecas = (na, nb)
        self.interface = interface
        Ea = _excitation_tables(ncas, na)
        Eb = _excitation_tables(ncas, nb)
        Da, Db = Ea.shape[-1], Eb.shape[-1]
        self.dim = Da * Db
        Ia, Ib = np.eye(Da), np.eye(Db)
        E = np.empty((ncas, ncas, self.dim, self.dim))
        for p in range(ncas):
            for q in range(ncas):
                E[p, q] = np.kron(Ea[p, q], Ib) + np.kron(Ia, Eb[p, q])
        self._E = torch.as_tensor(E)
        rng = np.random.default_rng(4200 + seed)
        gens = rng.standard_normal((n_theta, self.dim, self.dim))
        self._gens = torch.as_tensor(gens - gens.transpose(0, 2, 1))
        psi0 = np.zeros(self.dim)
        psi0[0] = 1.0                                   # lowest strings = aufbau determinant
        self._psi0 = torch.as_tensor(psi0)
        self.theta_shape = (n_theta,)

    def init_zeros(self):
        return torch.zeros(self.theta_shape, dtype=torch.float64)

    def state(self, theta):
        theta = torch.as_tensor(theta, dtype=torch.float64)
        gen = torch.einsum('k,kij->ij', theta.reshape(-1), self._gens)
        return torch.linalg.matrix_exp(gen) @ self._psi0

    def get_rdms_from_state(self, psi):
        n = self.ncas
        w = torch.einsum('pqij,j->pqi', self._E, psi)              # E_pq |psi>
        one = torch.einsum('i,pqi->pq', psi, w)
        two = torch.einsum('qpi,rsi->pqrs', w, w)                   # <E_qp psi | E_rs psi>
        eye = torch.eye(n, dtype=torch.float64)
        two = two - torch.einsum('qr,ps->pqrs', eye, one)
        return one, two

    def get_rdms(self, theta, restricted=True):
        return self.get_rdms_from_state(self.state(theta))
