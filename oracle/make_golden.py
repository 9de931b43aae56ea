"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the
VERBATIM reference (``/root/reference`` behind ``oracle/ref_shim.py``).

Run in the build container only (the reference does not travel to the GPU box):

    python -m oracle.make_golden

Each fixture stores the inputs (or, for N >= 28, the seeds that regenerate them
through ``auto_oo_b200.synthetic`` plus input checksums) and the reference's
outputs: rotation U, C', c0/c1/c2, E, packed gradient, Hessian matrix, and for
small N the transformed integrals.  The script also prints the oracle-vs-
reference differences so a regeneration doubles as an oracle check.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from auto_oo_b200.synthetic import SyntheticMol, random_rdms, random_kappa, CIVectorCircuit  # noqa: E402
from oracle import oo_oracle as orc                                                          # noqa: E402
from oracle.ref_shim import load_reference                                                   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name, nao, nelec, ncas, nelecas, freeze_active, kappa_scale, rdm kind, store_inputs
CASES = [
    ("n7_cas44",        7, 10, 4, 4, False, 0.05, "random", True),
    ("n7_cas44_frozen", 7, 10, 4, 4, True,  0.05, "ci",     True),
    ("n8_nocore",       8,  4, 3, 4, False, 0.05, "ci",     True),
    ("n11_cas43",      11, 12, 3, 4, False, 0.30, "random", True),
    ("n13_cas22",      13, 16, 2, 2, True,  0.05, "ci",     True),
    ("n13_bigkappa",   13, 16, 4, 4, False, 1.00, "random", True),
    ("n28_cas66",      28, 14, 6, 6, False, 0.05, "random", False),
    ("n34_cas44",      34, 16, 4, 4, True,  0.05, "random", False),
    ("n43_cas34",      43, 16, 3, 4, False, 0.05, "random", False),
]


def case_inputs(name, nao, nelec, ncas, nelecas, freeze, kscale, rdm_kind, seed):
    mol = SyntheticMol(nao, nelec, seed=seed)
    if rdm_kind == "ci":
        circ = CIVectorCircuit(ncas, nelecas, n_theta=3, seed=seed)
        theta = torch.tensor([0.31, -0.12, 0.07], dtype=torch.float64)
        one, two = circ.get_rdms(theta)
    else:
        one, two = random_rdms(ncas, nelecas, seed=seed)
    return mol, one.detach(), two.detach()


def reference_case(ref, mol, nelec, ncas, nelecas, freeze, C_oao, kappa_of, one, two, store, seed=0,
                   trajectory=False):
    """Run the verbatim reference on one problem; returns ``(fixture dict, oracle-vs-reference diffs)``.
    ``kappa_of(n_kappa)`` supplies the rotation once the reference has counted the parameters."""
    nao = mol.nao
    oo = ref.oo_energy.OO_energy(mol, ncas, nelecas, oao_mo_coeff=C_oao,
                                 freeze_active=freeze, interface='torch')
    kappa = kappa_of(oo.n_kappa)
    U = oo.kappa_to_mo_coeff(kappa)
    Cp = oo.mo_coeff @ U
    c0, c1, c2 = oo.get_active_integrals(Cp)
    E = oo.energy_from_kappa(kappa, one, two)
    G = oo.kappa_matrix_to_vector(oo.analytic_gradient(one, two, mo_coeff=Cp))
    H = oo.full_hessian_to_matrix(oo.analytic_hessian(one, two, mo_coeff=Cp))
    E0 = oo.energy_from_mo_coeff(oo.mo_coeff, one, two)
    G0 = oo.kappa_matrix_to_vector(oo.analytic_gradient(one, two))

    out = dict(
        shape=np.array([nao, nelec, ncas, nelecas, int(freeze)]), seed=np.array(seed),
        kappa=kappa.numpy(), one_rdm=one.numpy(), two_rdm=two.numpy(),
        params_idx=np.asarray(oo.params_idx), U=U.numpy(), mo_coeff_rot=Cp.numpy(),
        c0=np.asarray(float(c0)), c1=c1.numpy(), c2=c2.numpy(),
        E=np.asarray(E.item()), G=G.numpy(), H=H.numpy(),
        E0=np.asarray(E0.item()), G0=G0.numpy(),
        checksum=np.array([float(np.sum(mol.int1e_ao)), float(np.sum(mol.int2e_ao)),
                           float(np.sum(mol.oao_coeff)), float(np.sum(np.asarray(C_oao)))]),
    )
    if store:
        out.update(int1e_ao=mol.int1e_ao, int2e_ao=mol.int2e_ao, overlap=mol.overlap,
                   oao_coeff=mol.oao_coeff, oao_mo_coeff=np.asarray(C_oao), nuc=np.asarray(mol.nuc),
                   int1e_mo=ref.oo_energy.int1e_transform(oo.int1e_ao, Cp).numpy(),
                   int2e_mo=ref.oo_energy.int2e_transform(oo.int2e_ao, Cp).numpy())
    if trajectory:
        # orbital-only Newton-Raphson trajectory (oo_energy.py:426-474), fresh object
        oo2 = ref.oo_energy.OO_energy(mol, ncas, nelecas, oao_mo_coeff=C_oao,
                                      freeze_active=freeze, interface='torch')
        with contextlib.redirect_stdout(io.StringIO()):
            traj = oo2.orbital_optimization(one, two, conv_tol=1e-10, max_iterations=8, verbose=0)
        out.update(nr_energies=np.asarray(traj), nr_oao_mo_coeff=oo2.oao_mo_coeff.numpy())

    # oracle cross-check
    prob = orc.OracleProblem(mol.int1e_ao, mol.int2e_ao, mol.oao_coeff, C_oao, mol.nuc,
                             nelec, ncas, nelecas, freeze)
    eo, go, ho = prob.evaluate(one, two, kappa)
    hi = prob.hessian(one, two, kappa, ispace=True)
    d = (abs(eo.item() - E.item()), (go - G).abs().max().item(), (ho - H).abs().max().item(),
         (hi - H).abs().max().item())
    return out, d, oo.n_kappa


def main():
    ref = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_default_dtype(torch.float64)
    worst = 0.0
    for seed, (name, nao, nelec, ncas, nelecas, freeze, kscale, rdm_kind, store) in enumerate(CASES):
        mol, one, two = case_inputs(name, nao, nelec, ncas, nelecas, freeze, kscale, rdm_kind, seed)
        out, d, nk = reference_case(ref, mol, nelec, ncas, nelecas, freeze, mol.random_oao_mo_coeff,
                                    lambda n: random_kappa(n, seed=seed, scale=kscale), one, two, store,
                                    seed=seed, trajectory=nao <= 13)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        worst = max(worst, *d)
        print(f"{name:18s} N={nao:3d} nk={nk:4d} E={float(out['E']):+.10f} "
              f"|dE|={d[0]:.1e} |dG|={d[1]:.1e} |dH|={d[2]:.1e} |dH_ispace|={d[3]:.1e}")

    # general (non-symmetric) 4-index transform with four different matrices
    rng = np.random.default_rng(5)
    M = rng.standard_normal((6, 6, 6, 6))
    Cs = [rng.standard_normal((6, 6)) for _ in range(4)]
    Mt = ref.oo_energy.general_4index_transform(torch.as_tensor(M), *[torch.as_tensor(c) for c in Cs])
    np.savez_compressed(os.path.join(GOLDEN, "general_4index_n6.npz"), M=M, C0=Cs[0], C1=Cs[1],
                        C2=Cs[2], C3=Cs[3], out=Mt.numpy())
    d = (orc.transform_4index(M, *Cs) - Mt).abs().max().item()
    worst = max(worst, d)
    print(f"general_4index_n6  |d|={d:.1e}")

    # NewtonStep known answers (newton_raphson.py:78-211) on a small quartic objective
    a = torch.as_tensor(rng.standard_normal((5, 5)))
    a = a + a.T
    x0 = torch.as_tensor(rng.standard_normal(5)) * 0.3

    def f(x):
        return 0.5 * x @ a @ x + 0.25 * torch.sum(x ** 4) + torch.sum(x)

    g0 = torch.autograd.functional.jacobian(f, x0)
    h0 = torch.autograd.functional.hessian(f, x0)
    opt = ref.newton_raphson.NewtonStep(verbose=0)
    dp, lam = opt.newton_step(g0, h0)
    newx, lam2 = opt.damped_newton_step(f, (x0,), g0, h0)
    np.savez_compressed(os.path.join(GOLDEN, "newton_step.npz"), a=a.numpy(), x0=x0.numpy(),
                        grad=g0.numpy(), hess=h0.numpy(), dp=dp.numpy(), lowest=np.asarray(lam),
                        new_x=newx.numpy())
    print("newton_step lowest eig", lam, "| worst oracle-vs-reference diff", worst)


if __name__ == "__main__":
    main()
