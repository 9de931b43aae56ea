"""Interchange format for molecular data produced OUTSIDE this package (e.g. by the reference's
``Moldata_pyscf`` on a machine that has PySCF): one ``.npz`` holding what ``OO_energy.__init__`` reads
(reference ``oo_energy.py:143-165``) plus optional orbitals and optimisation state.

    save_problem("ch2nh.npz", mol, oao_mo_coeff=C, theta=theta)      # where PySCF exists
    mol, extras = load_problem("ch2nh.npz")                          # anywhere
    oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=extras["oao_mo_coeff"])
"""
from __future__ import annotations

import numpy as np

FORMAT_VERSION = 1
_REQUIRED = ("int1e_ao", "int2e_ao", "overlap", "oao_coeff", "nuc", "nelectron")


class ArrayMol:
    """Duck-typed ``Moldata_pyscf`` built from arrays (``moldata_pyscf.py:19-56``): attributes
    ``int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nao, nelectron`` and ``get_active_space_idx``."""

    def __init__(self, int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nelectron):
        self.int1e_ao = np.asarray(int1e_ao, dtype=np.float64)
        self.int2e_ao = np.asarray(int2e_ao, dtype=np.float64)
        self.overlap = np.asarray(overlap, dtype=np.float64)
        self.oao_coeff = np.asarray(oao_coeff, dtype=np.float64)
        self.nuc = float(nuc)
        self.nelectron = int(nelectron)
        self.nao = self.int1e_ao.shape[0]
        if self.int2e_ao.shape != (self.nao,) * 4 or self.overlap.shape != (self.nao,) * 2:
            raise ValueError("inconsistent array shapes for an AO basis of size %d" % self.nao)

    def get_active_space_idx(self, ncas, nelecas):
        """Same rule (and error) as ``moldata_pyscf.py:42-56``."""
        nelecore = self.nelectron - nelecas
        if nelecore % 2 == 1:
            raise ValueError('odd number of core electrons')
        occ_idx = np.arange(nelecore // 2)
        act_idx = (occ_idx[-1] + 1 + np.arange(ncas)) if len(occ_idx) > 0 else np.arange(ncas)
        virt_idx = np.arange(act_idx[-1] + 1, self.nao)
        return occ_idx, act_idx, virt_idx


def save_problem(path, mol, nelectron=None, **extras):
    """Write ``mol``'s integrals (and any extra arrays: ``oao_mo_coeff``, ``theta``, RDMs, ...)."""
    ne = nelectron if nelectron is not None else getattr(mol, "nelectron", None)
    if ne is None:
        raise ValueError("nelectron is required (mol has no .nelectron)")
    data = dict(format_version=np.asarray(FORMAT_VERSION), int1e_ao=np.asarray(mol.int1e_ao),
                int2e_ao=np.asarray(mol.int2e_ao), overlap=np.asarray(mol.overlap),
                oao_coeff=np.asarray(mol.oao_coeff), nuc=np.asarray(float(mol.nuc)), nelectron=np.asarray(int(ne)))
    for k, v in extras.items():
        if k in data:
            raise ValueError(f"extra array name {k!r} collides with a required field")
        data[k] = np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
    np.savez_compressed(path, **data)


def load_problem(path):
    """Returns ``(ArrayMol, extras dict)``."""
    with np.load(path) as d:
        missing = [k for k in _REQUIRED if k not in d.files]
        if missing:
            raise ValueError(f"{path}: missing fields {missing}")
        if int(d["format_version"]) != FORMAT_VERSION:
            raise ValueError(f"{path}: unsupported format version {int(d['format_version'])}")
        mol = ArrayMol(d["int1e_ao"], d["int2e_ao"], d["overlap"], d["oao_coeff"], float(d["nuc"]),
                       int(d["nelectron"]))
        extras = {k: d[k] for k in d.files if k not in _REQUIRED and k != "format_version"}
    return mol, extras
