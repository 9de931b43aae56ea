"""Shared fixtures: golden cases produced by the verbatim reference
(``oracle/make_golden.py``) and the matching synthetic inputs."""
import os

import numpy as np
import torch

from auto_oo_b200.synthetic import SyntheticMol

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# real molecules (STO-3G integrals rebuilt by oracle/gto_sto3g.py, outputs by the verbatim reference,
# oracle/make_molecular_golden.py): formaldimine at the reference's two test geometries, water (config 1)
MOL_CASES = ["mol_ch2nh_sto3g_cas22", "mol_ch2nh_sto3g_cas44", "mol_h2o_sto3g_cas44"]
SMALL_CASES = ["n7_cas44", "n7_cas44_frozen", "n8_nocore", "n11_cas43", "n13_cas22", "n13_bigkappa"] + MOL_CASES
SEEDED_CASES = ["n28_cas66", "n34_cas44", "n43_cas34"]
ALL_CASES = SMALL_CASES + SEEDED_CASES

# energy 1e-10 Ha, gradient / Hessian elements 1e-8 (BASELINE.json north_star)
TOL_E = 1e-10
TOL_GH = 1e-8


class Case:
    """Inputs and reference outputs of one golden fixture."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.ref = {k: d[k] for k in d.files}
        nao, nelec, ncas, nelecas, freeze = (int(x) for x in d["shape"])
        self.nao, self.nelec, self.ncas, self.nelecas, self.freeze = nao, nelec, ncas, nelecas, bool(freeze)
        if "int2e_ao" in d.files:
            self.int1e_ao, self.int2e_ao = d["int1e_ao"], d["int2e_ao"]
            self.overlap, self.oao_coeff = d["overlap"], d["oao_coeff"]
            self.oao_mo_coeff, self.nuc = d["oao_mo_coeff"], float(d["nuc"])
        else:
            mol = SyntheticMol(nao, nelec, seed=int(d["seed"]))
            self.int1e_ao, self.int2e_ao = mol.int1e_ao, mol.int2e_ao
            self.overlap, self.oao_coeff = mol.overlap, mol.oao_coeff
            self.oao_mo_coeff, self.nuc = mol.random_oao_mo_coeff, mol.nuc
            chk = np.array([np.sum(self.int1e_ao), np.sum(self.int2e_ao), np.sum(self.oao_coeff),
                            np.sum(self.oao_mo_coeff)])
            assert np.allclose(chk, d["checksum"], rtol=1e-12, atol=1e-9), \
                "seeded inputs no longer reproduce the fixture's inputs"
        self.kappa = torch.as_tensor(d["kappa"])
        self.one_rdm = torch.as_tensor(d["one_rdm"])
        self.two_rdm = torch.as_tensor(d["two_rdm"])

    def mol(self):
        """Duck-typed Moldata_pyscf (moldata_pyscf.py:19-56)."""
        from oracle.ref_shim import FakeMol
        return FakeMol(self.int1e_ao, self.int2e_ao, self.overlap, self.oao_coeff, self.nuc, self.nelec)

    def oracle(self):
        from oracle import oo_oracle as orc
        return orc.OracleProblem(self.int1e_ao, self.int2e_ao, self.oao_coeff, self.oao_mo_coeff,
                                 self.nuc, self.nelec, self.ncas, self.nelecas, self.freeze)


_cache = {}


def load_case(name):
    if name not in _cache:
        _cache[name] = Case(name)
    return _cache[name]
