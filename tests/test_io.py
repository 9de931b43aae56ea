"""CPU: the npz interchange format and the array-backed Moldata stand-in."""
import numpy as np
import pytest

from auto_oo_b200.io import ArrayMol, load_problem, save_problem
from auto_oo_b200.synthetic import SyntheticMol


def test_roundtrip(tmp_path):
    mol = SyntheticMol(7, 10, seed=1)
    C = mol.random_oao_mo_coeff
    path = tmp_path / "p.npz"
    save_problem(path, mol, oao_mo_coeff=C, theta=np.array([0.1, 0.2]))
    m2, extras = load_problem(path)
    assert m2.nao == 7 and m2.nelectron == 10 and m2.nuc == mol.nuc
    for name in ("int1e_ao", "int2e_ao", "overlap", "oao_coeff"):
        assert np.array_equal(getattr(m2, name), np.asarray(getattr(mol, name)))
    assert np.array_equal(extras["oao_mo_coeff"], C) and np.array_equal(extras["theta"], [0.1, 0.2])
    occ, act, virt = m2.get_active_space_idx(4, 4)
    assert list(occ) == [0, 1, 2] and list(act) == [3, 4, 5, 6] and len(virt) == 0


def test_errors(tmp_path):
    mol = SyntheticMol(4, 4, seed=0)
    with pytest.raises(ValueError):
        ArrayMol(mol.int1e_ao, mol.int2e_ao[:3], mol.overlap, mol.oao_coeff, 0.0, 4)
    with pytest.raises(ValueError):
        ArrayMol(mol.int1e_ao, mol.int2e_ao, mol.overlap, mol.oao_coeff, 0.0, 5).get_active_space_idx(2, 2)
    with pytest.raises(ValueError):
        save_problem(tmp_path / "x.npz", mol, nuc=1.0)
    np.savez(tmp_path / "bad.npz", int1e_ao=np.eye(2))
    with pytest.raises(ValueError):
        load_problem(tmp_path / "bad.npz")
