"""GPU: "the existing Newton-Raphson and Berry-phase drivers run unchanged" (BASELINE.json north_star).

The reference's own driver code -- ``OO_pqc`` (``oo_pqc.py:30-207``: ``full_gradient``, ``full_hessian``,
``full_optimization``) and ``NewtonStep`` (``utils/newton_raphson.py:16-224``) -- is executed VERBATIM from the
unmodified install under ``baseline/_ref`` (``oracle/ref_shim.py``; ``/root/reference`` in the build
container), once on top of the reference's own ``OO_energy`` (CPU, all reference) and once with
``auto_oo.oo_energy`` bound to ``auto_oo_b200.oo_energy``, i.e. the reference's ``class OO_pqc(OO_energy)``
derived from the CUDA implementation.  CPU float64 tensors go in and come out on both sides; the
trajectories must agree.  The Berry-phase loop is the notebook's cell 22
(``examples/Tutorial_Berry_phase.ipynb``) restated around the same verbatim classes.

PennyLane circuits are out of scope: ``pqc`` is the CI-vector stand-in with the two members the drivers use
(``get_rdms(theta)``, ``theta_shape``)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle.ref_shim import FakeMol, load_reference, load_reference_oo_pqc, reference_available

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(), reason="no reference install (baseline/_ref)")]


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _driver_classes():
    import auto_oo_b200.oo_energy as cuda_oo_energy
    ref = load_reference()
    return (load_reference_oo_pqc().OO_pqc, load_reference_oo_pqc(cuda_oo_energy).OO_pqc,
            ref.newton_raphson.NewtonStep)


def _problem(nao, nelec, ncas, nelecas, seed):
    from auto_oo_b200.synthetic import SyntheticMol, CIVectorCircuit
    mol = SyntheticMol(nao, nelec, seed=seed)
    pqc = CIVectorCircuit(ncas, nelecas, n_theta=3, seed=seed)
    return mol, pqc


@pytest.mark.parametrize("shape", [(11, 12, 3, 4, True), (13, 16, 2, 2, False)])
def test_reference_full_optimization_runs_unchanged_on_the_cuda_oo_energy(shape):
    nao, nelec, ncas, nelecas, freeze = shape
    RefPqc, CudaPqc, _ = _driver_classes()
    import auto_oo_b200.oo_energy as cuda_oo_energy
    assert issubclass(CudaPqc, cuda_oo_energy.OO_energy) and not issubclass(RefPqc, cuda_oo_energy.OO_energy)
    mol, pqc = _problem(nao, nelec, ncas, nelecas, seed=3)
    theta0 = torch.tensor([0.1, -0.05, 0.02], dtype=torch.float64)
    runs = []
    for cls in (RefPqc, CudaPqc):
        oo = cls(pqc, mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, freeze_active=freeze)
        runs.append(_quiet(oo.full_optimization, theta0.clone(), max_iterations=4, verbose=0))
    (e_r, th_r, _, c_r, eig_r), (e_c, th_c, _, c_c, eig_c) = runs
    assert len(e_r) == len(e_c) == 4
    for n in range(4):
        assert e_c[n] == pytest.approx(e_r[n], abs=1e-8)
        assert eig_c[n] == pytest.approx(eig_r[n], abs=1e-7)
        assert torch.allclose(th_c[n], th_r[n], rtol=0, atol=1e-7)
        assert th_c[n].device.type == "cpu" and c_c[n].device.type == "cpu"
        assert torch.allclose(torch.as_tensor(c_c[n]), torch.as_tensor(c_r[n]), rtol=0, atol=1e-7)


def _berry_loop(OO_pqc, NewtonStep, pqc, mols, ncas, nelecas, theta0, oao0):
    """examples/Tutorial_Berry_phase.ipynb cell 22, statement for statement (PySCF / CASSCF lines dropped)."""
    opt = NewtonStep(verbose=0)
    curr_theta, curr_oao_mo_coeff = theta0, oao0
    theta_l, oao_mo_coeff_l, energy_l, hess_eig_l = [], [], [], []
    for step, mol in enumerate(mols):
        if step == 0:
            continue
        oo_pqc = OO_pqc(pqc, mol, ncas, nelecas, freeze_active=True, oao_mo_coeff=oao0)
        oo_pqc.oao_mo_coeff = curr_oao_mo_coeff
        kappa = torch.zeros(oo_pqc.n_kappa, dtype=torch.float64)
        gradient = oo_pqc.full_gradient(curr_theta)
        hessian = oo_pqc.full_hessian(curr_theta)
        new_parameters, hess_eig = opt.damped_newton_step(
            oo_pqc.energy_from_parameters, (curr_theta, kappa), gradient, hessian)
        curr_theta = new_parameters[0]
        kappa = new_parameters[1]
        curr_oao_mo_coeff = curr_oao_mo_coeff @ oo_pqc.kappa_to_mo_coeff(kappa)
        oo_pqc.oao_mo_coeff = curr_oao_mo_coeff.detach().clone()
        energy = oo_pqc.energy_from_parameters(curr_theta).item()
        theta_l.append(curr_theta.detach().clone())
        oao_mo_coeff_l.append(curr_oao_mo_coeff.detach().clone())
        energy_l.append(energy)
        hess_eig_l.append(hess_eig)
    return theta_l, oao_mo_coeff_l, energy_l, hess_eig_l


def test_berry_phase_loop_cell_22_on_the_cuda_oo_energy():
    """Five 'geometries' = the integrals of one synthetic molecule moved a few percent towards another."""
    from auto_oo_b200.synthetic import SyntheticMol, CIVectorCircuit
    RefPqc, CudaPqc, NewtonStep = _driver_classes()
    nao, nelec, ncas, nelecas = 13, 16, 2, 2
    a, b = SyntheticMol(nao, nelec, seed=21), SyntheticMol(nao, nelec, seed=22)
    mols = []
    for k in range(5):
        t = 0.02 * k
        mix = lambda x, y: (1 - t) * np.asarray(x) + t * np.asarray(y)
        S = mix(a.overlap, b.overlap)
        w, v = np.linalg.eigh(S)
        mols.append(FakeMol(mix(a.int1e_ao, b.int1e_ao), mix(a.int2e_ao, b.int2e_ao), S, (v * w ** -0.5) @ v.T,
                            a.nuc + 0.1 * k, nelec))
    pqc = CIVectorCircuit(ncas, nelecas, n_theta=3, seed=5)
    theta0 = torch.tensor([0.05, -0.02, 0.01], dtype=torch.float64)
    oao0 = torch.as_tensor(a.random_oao_mo_coeff)
    ref = _quiet(_berry_loop, RefPqc, NewtonStep, pqc, mols, ncas, nelecas, theta0, oao0)
    cud = _quiet(_berry_loop, CudaPqc, NewtonStep, pqc, mols, ncas, nelecas, theta0, oao0)
    for (th_r, c_r, e_r, l_r), (th_c, c_c, e_c, l_c) in zip(zip(*ref), zip(*cud)):
        assert e_c == pytest.approx(e_r, abs=1e-8)
        assert l_c == pytest.approx(l_r, abs=1e-7)
        assert torch.allclose(th_c, th_r, rtol=0, atol=1e-7)
        assert torch.allclose(c_c, c_r, rtol=0, atol=1e-7)


def test_reference_newton_step_with_the_cuda_orbital_gradient_and_hessian():
    """``NewtonStep.damped_newton_step`` (newton_raphson.py:194-211), verbatim, driven the way
    ``OO_energy.orbital_optimization`` drives it (oo_energy.py:446-458) with CPU tensors from the CUDA path."""
    from auto_oo_b200 import OO_energy
    from auto_oo_b200.synthetic import SyntheticMol, random_rdms
    from functools import partial
    ref = load_reference()
    mol = SyntheticMol(28, 14, seed=6)
    one, two = random_rdms(6, 6, seed=6)
    args = (mol, 6, 6)
    oo_c = OO_energy(*args, oao_mo_coeff=mol.random_oao_mo_coeff)
    oo_r = ref.oo_energy.OO_energy(*args, oao_mo_coeff=mol.random_oao_mo_coeff)
    out = []
    for oo in (oo_r, oo_c):
        opt = ref.newton_raphson.NewtonStep(verbose=0)
        kappa = torch.zeros(oo.n_kappa, dtype=torch.float64)
        g = oo.kappa_matrix_to_vector(oo.analytic_gradient(one, two))
        h = oo.full_hessian_to_matrix(oo.analytic_hessian(one, two))
        assert g.device.type == "cpu" and h.device.type == "cpu"
        newk, lam = _quiet(opt.damped_newton_step, partial(oo.energy_from_kappa, one_rdm=one, two_rdm=two),
                           (kappa,), g, h)
        out.append((newk, lam, oo.energy_from_kappa(newk, one, two).item()))
    (k_r, l_r, e_r), (k_c, l_c, e_c) = out
    assert torch.allclose(k_c, k_r, rtol=0, atol=1e-8)
    assert l_c == pytest.approx(l_r, abs=1e-8) and e_c == pytest.approx(e_r, abs=1e-9)
