"""TEST INFRASTRUCTURE ONLY -- golden fixtures at the two full-size BASELINE configurations.

    python -m oracle.make_golden_large n114      # config 4 shape: 114 AOs, CAS(6,6)   (~3 min, ~20 GB)
    python -m oracle.make_golden_large n256      # config 5:       256 AOs, CAS(12,12) (~10 min, ~50 GB)

Run in the build container only.  Inputs are the seeded ``SyntheticMol`` / ``random_rdms`` /
``random_kappa`` objects drawn on the CPU (the GPU tests regenerate them with ``rng_device="cpu"``
and check the stored input checksums); the fixtures hold outputs only:

* ``n114_cas66.npz`` -- produced by the VERBATIM reference (``oracle/ref_shim.py``): ``energy_from_kappa``,
  ``kappa_matrix_to_vector(analytic_gradient(mo_coeff=C'))`` and
  ``full_hessian_to_matrix(analytic_hessian(mo_coeff=C'))`` (oo_energy.py:199-202, :404-424).  The
  2283 x 2283 Hessian is kept as its diagonal, its products with 8 seeded probe vectors, 20 000 seeded
  sampled elements and its Frobenius norm.  The same run checks ``oracle/class_oracle.py`` and
  ``oracle/oo_oracle.py`` against the reference at this size (differences are stored in the fixture).
* ``n256_cas1212.npz`` -- the reference cannot run at N = 256 (a dozen N^4 tensors, 400 GB); produced by
  ``oracle/class_oracle.py`` (pinned against the verbatim reference at N = 7 ... 114), same contents
  for the 9778 x 9778 Hessian.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa   # noqa: E402
from oracle import class_oracle as corc                                                     # noqa: E402
from oracle import oo_oracle as orc                                                         # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 11                      # tests/test_gpu_full_size.py builds its problems from the same seed
N_PROBES, N_SAMPLES = 8, 20000


def probes_and_samples(nk):
    rng = np.random.default_rng(1234 + nk)
    V = rng.standard_normal((N_PROBES, nk))
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    si = rng.integers(0, nk, size=N_SAMPLES)
    sj = rng.integers(0, nk, size=N_SAMPLES)
    return V, si, sj


def hessian_digest(H, nk):
    H = H.numpy() if torch.is_tensor(H) else H
    V, si, sj = probes_and_samples(nk)
    return dict(H_diag=np.diagonal(H).copy(), H_probes=V, H_times_probes=V @ H.T, H_sample_i=si,
                H_sample_j=sj, H_samples=H[si, sj].copy(), H_fro=np.asarray(np.linalg.norm(H)),
                H_asym=np.asarray(np.abs(H - H.T).max()))


def inputs(workload, build_eri=True):
    nao, nelec, ncas, nelecas = CONFIG_SHAPES[workload]
    mol = SyntheticMol(nao, nelec, seed=SEED, build_eri=build_eri)
    one, two = random_rdms(ncas, nelecas, seed=SEED)
    return mol, one, two, (nao, nelec, ncas, nelecas)


def checksum(mol):
    """Sums of the seeded inputs; the ERI entry is the sum of the density-fitting factor (the GPU test forms
    B B^T on the device and compares its own sum of g with ``g_sum`` to 1e-10 relative)."""
    return np.array([float(mol._int1e.sum()), float(mol._B.sum()), float(mol._oao.sum()),
                     float(mol._oao_mo.sum()), float((mol._B ** 2).sum())])


def make_n114():
    from oracle.ref_shim import load_reference
    ref = load_reference()
    torch.set_default_dtype(torch.float64)
    mol, one, two, (nao, nelec, ncas, nelecas) = inputs("c6h6_ccpvdz_cas66")
    t0 = time.time()
    oo = ref.oo_energy.OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff,
                                 freeze_active=False, interface='torch')
    kappa = random_kappa(oo.n_kappa, seed=SEED, batch=2)[0]
    Cp = oo.mo_coeff @ oo.kappa_to_mo_coeff(kappa)
    E = oo.energy_from_kappa(kappa, one, two)
    print(f"reference E {E.item():+.12f}  ({time.time() - t0:.0f} s)", flush=True)
    G = oo.kappa_matrix_to_vector(oo.analytic_gradient(one, two, mo_coeff=Cp))
    print(f"reference G done ({time.time() - t0:.0f} s)", flush=True)
    H = oo.full_hessian_to_matrix(oo.analytic_hessian(one, two, mo_coeff=Cp))
    print(f"reference H done ({time.time() - t0:.0f} s)", flush=True)
    c0, c1, c2 = oo.get_active_integrals(Cp)
    nk = oo.n_kappa

    # the two oracles against the verbatim reference at this size
    cp = corc.ClassProblem(mol.int1e_ao, mol.int2e_ao, mol.oao_coeff, mol.random_oao_mo_coeff, mol.nuc,
                           nelec, ncas, nelecas, False)
    Ec, Gc, Hc = cp.evaluate(one, two, kappa)
    d_class = [abs(Ec.item() - E.item()), (Gc - G).abs().max().item(), (Hc - H).abs().max().item()]
    print("class_oracle vs reference |dE|, |dG|, |dH|:", d_class, flush=True)
    op = orc.OracleProblem(mol.int1e_ao, mol.int2e_ao, mol.oao_coeff, mol.random_oao_mo_coeff, mol.nuc,
                           nelec, ncas, nelecas, False)
    Eo = op.energy(one, two, kappa)
    Go = op.gradient(one, two, kappa)
    Ho = op.hessian(one, two, kappa, ispace=True)
    d_orc = [abs(Eo.item() - E.item()), (Go - G).abs().max().item(), (Ho - H).abs().max().item()]
    print("oo_oracle (I-space Hessian) vs reference:", d_orc, flush=True)

    out = dict(shape=np.array([nao, nelec, ncas, nelecas, 0]), seed=np.array(SEED), source=np.array("verbatim reference"),
               kappa=kappa.numpy(), E=np.asarray(E.item()), G=G.numpy(), c0=np.asarray(float(c0)), c1=c1.numpy(),
               c2=c2.numpy(), checksum=checksum(mol), g_sum=np.asarray(float(mol._int2e.sum())),
               class_oracle_vs_reference=np.array(d_class), oo_oracle_vs_reference=np.array(d_orc),
               **hessian_digest(H, nk))
    np.savez_compressed(os.path.join(GOLDEN, "n114_cas66.npz"), **out)
    print("wrote n114_cas66.npz", flush=True)


def make_n256():
    torch.set_default_dtype(torch.float64)
    t0 = time.time()
    mol, one, two, (nao, nelec, ncas, nelecas) = inputs("synthetic_n256_cas1212", build_eri=False)
    chk = checksum(mol)
    g = mol.build_eri()                                    # 34.4 GB on the host
    mol._B = None
    g_sum = float(g.sum())
    print(f"g built ({time.time() - t0:.0f} s)", flush=True)
    occ, act, virt = orc.active_space_idx(nao, nelec, ncas, nelecas)
    params_idx = orc.non_redundant_indices(occ, act, virt, False)
    nk = len(params_idx)
    kappa = random_kappa(nk, seed=SEED, batch=2)[0]
    C = mol._oao @ mol._oao_mo @ orc.rotation_from_kappa(kappa, params_idx, nao)
    h, J, K = corc.class_integrals(mol._int1e, g, C, len(occ) + len(act))
    del g
    mol._int2e = None
    print(f"class integrals ({time.time() - t0:.0f} s)", flush=True)
    ev = corc.ClassEvaluation(h, J, K, mol.nuc, len(occ), len(act), params_idx)
    E = ev.energy(one, two)
    G = ev.gradient(one, two)
    c0, c1, c2 = ev.hamiltonian()
    print(f"E {E.item():+.12f}, G done ({time.time() - t0:.0f} s)", flush=True)
    H = ev.hessian(one, two)
    print(f"H done ({time.time() - t0:.0f} s)", flush=True)
    out = dict(shape=np.array([nao, nelec, ncas, nelecas, 0]), seed=np.array(SEED),
               source=np.array("oracle/class_oracle.py (pinned on the verbatim reference up to N=114)"),
               kappa=kappa.numpy(), E=np.asarray(E.item()), G=G.numpy(), c0=np.asarray(float(c0)), c1=c1.numpy(),
               c2=c2.numpy(), checksum=chk, g_sum=np.asarray(g_sum), **hessian_digest(H, nk))
    np.savez_compressed(os.path.join(GOLDEN, "n256_cas1212.npz"), **out)
    print("wrote n256_cas1212.npz", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["n114"]
    torch.set_num_threads(os.cpu_count() or 1)
    if "n114" in which:
        make_n114()
    if "n256" in which:
        make_n256()
