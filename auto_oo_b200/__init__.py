"""B200-native orbital-optimisation hot path of ``auto_oo`` (same public names as the
reference's ``auto_oo/__init__.py:3-27`` for the path in scope)."""
from .oo_energy import (OO_energy, OO_energy_geometries, OrbitalHessian, PendingEvaluation, unpack_hessian, mo_ao_to_mo_oao, int1e_transform, int2e_transform,
                        general_4index_transform, uniform_4index_transform, vector_to_skew_symmetric,
                        skew_symmetric_to_vector, non_redundant_indices)
from .oo_pqc import OO_pqc
from .rdm import StatevectorRDM, StatevectorCircuit
from .utils.newton_raphson import NewtonStep

__all__ = ["OO_energy", "OO_energy_geometries", "OO_pqc", "OrbitalHessian", "PendingEvaluation", "unpack_hessian", "NewtonStep", "StatevectorRDM", "StatevectorCircuit", "mo_ao_to_mo_oao", "int1e_transform",
           "int2e_transform", "general_4index_transform", "uniform_4index_transform",
           "vector_to_skew_symmetric", "skew_symmetric_to_vector", "non_redundant_indices"]
