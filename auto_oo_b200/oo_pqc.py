"""``OO_pqc``: the hybrid cost function that feeds circuit RDMs into the orbital hot path
(reference ``src/auto_oo/oo_pqc.py:30-207``).  The circuit object is duck-typed --
``get_rdms(theta) -> (one_rdm, two_rdm)`` and ``theta_shape`` -- so a PennyLane
``Parameterized_circuit`` or any other differentiable RDM provider works.

Derivatives w.r.t. the circuit parameters come from ``torch.autograd.functional`` exactly as in the
reference; they flow through the hot path by the RDM-side adjoints of
:class:`auto_oo_b200.oo_energy.OO_energy` (energy: ``dE/dgamma = c1``, ``dE/dGamma = c2``; orbital
gradient: ``oo_fock_gradient_vjp_f64``)."""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd.functional import jacobian, hessian

from .oo_energy import OO_energy
from .utils.newton_raphson import NewtonStep


class OO_pqc(OO_energy):
    def __init__(self, pqc, mol, ncas, nelecas, oao_mo_coeff=None, freeze_active=False,
                 interface='torch', device=None):
        super().__init__(mol, ncas, nelecas, oao_mo_coeff=oao_mo_coeff, freeze_active=freeze_active,
                         interface=interface, device=device)
        self.pqc = pqc

    @property
    def _n_theta(self):
        return int(np.prod(self.pqc.theta_shape))

    def energy_from_parameters(self, theta, kappa=None):
        """Total energy at circuit parameters ``theta`` and rotation ``kappa`` (reference ``:64-84``)."""
        one_rdm, two_rdm = self.pqc.get_rdms(theta)
        if kappa is None:
            return self.energy_from_mo_coeff(self.mo_coeff, one_rdm, two_rdm)
        return self.energy_from_kappa(kappa, one_rdm, two_rdm)

    def circuit_gradient(self, theta):
        """Reference ``:86-93``."""
        return jacobian(self.energy_from_parameters, theta).reshape(self._n_theta)

    def orbital_gradient(self, theta):
        """Packed analytic orbital gradient (reference ``:97-101``)."""
        one_rdm, two_rdm = self.pqc.get_rdms(theta)
        return self.kappa_matrix_to_vector(self.analytic_gradient(one_rdm, two_rdm))

    def circuit_circuit_hessian(self, theta):
        """Reference ``:103-111``."""
        return hessian(self.energy_from_parameters, theta).reshape(self._n_theta, self._n_theta)

    def orbital_circuit_hessian(self, theta):
        """Mixed block: autograd of the analytic orbital gradient (reference ``:113-123``)."""
        return jacobian(self.orbital_gradient, theta).reshape(self.n_kappa, self._n_theta)

    def orbital_orbital_hessian(self, theta):
        """Reference ``:127-130``."""
        one_rdm, two_rdm = self.pqc.get_rdms(theta)
        return self.full_hessian_to_matrix(
            self.analytic_hessian(one_rdm.detach(), two_rdm.detach()))

    def full_gradient(self, theta):
        """Reference ``:132-134``."""
        return torch.cat((self.circuit_gradient(theta), self.orbital_gradient(theta).detach()))

    def full_hessian(self, theta):
        """Reference ``:136-148``."""
        h_cc = self.circuit_circuit_hessian(theta)
        h_oc = self.orbital_circuit_hessian(theta)
        h_oo = self.orbital_orbital_hessian(theta)
        return torch.cat((torch.cat((h_cc, h_oc.T), dim=1), torch.cat((h_oc, h_oo), dim=1)), dim=0)

    def full_circuit_hessian_to_matrix(self, full_circuit_hessian):
        return full_circuit_hessian.reshape(self._n_theta, self._n_theta)

    def full_optimization(self, theta_init, max_iterations=50, conv_tol=1e-10, verbose=0, flush=True,
                          **kwargs):
        """Joint Newton-Raphson optimisation of circuit parameters and orbitals (reference
        ``:155-207``; ``kappa_l`` holds the accepted rotations -- the reference appends ``theta``
        there by mistake, ``:189``)."""
        opt = NewtonStep(verbose=verbose, **kwargs)
        energy_init = self.energy_from_parameters(theta_init).item()
        if verbose is not None:
            print(f"iter = 000, energy = {energy_init:.12f}", flush=flush)
        theta_l, kappa_l, oao_mo_coeff_l, energy_l, hess_eig_l = [], [], [], [], []
        theta = theta_init
        for n in range(max_iterations):
            kappa = torch.zeros(self.n_kappa, dtype=theta_init.dtype, device=theta_init.device)
            grad = self.full_gradient(theta)
            hess = self.full_hessian(theta)
            new_theta_kappa, hess_eig = opt.damped_newton_step(
                self.energy_from_parameters, (theta, kappa), grad, hess)
            hess_eig_l.append(hess_eig)
            theta = new_theta_kappa[0].reshape(self.pqc.theta_shape)
            kappa = new_theta_kappa[1]
            theta_l.append(theta)
            kappa_l.append(kappa)
            self.oao_mo_coeff = self.get_transformed_mo(self.oao_mo_coeff, kappa)
            oao_mo_coeff_l.append(self.oao_mo_coeff)
            energy = self.energy_from_parameters(theta).item()
            energy_l.append(energy)
            if verbose is not None:
                print(f"iter = {n + 1:03}, energy = {energy:.12f}")
            if n > 1 and abs(energy_l[-1] - energy_l[-2]) < conv_tol:
                if verbose is not None:
                    print("optimization finished.")
                    print("E_fin =", energy_l[-1])
                break
        return energy_l, theta_l, kappa_l, oao_mo_coeff_l, hess_eig_l
