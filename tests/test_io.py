"""CPU: the npz interchange format and the array-backed Moldata stand-in."""
import numpy as np
import pytest

from auto_oo_b200.io import (ArrayMol, load_problem, load_trajectory, pack_eri_s8, save_problem, save_trajectory,
                             unpack_eri_s8)
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


def test_s8_packing_roundtrip_and_size(tmp_path):
    mol = SyntheticMol(9, 10, seed=3)
    g = np.asarray(mol.int2e_ao)
    packed = pack_eri_s8(g)
    P = 9 * 10 // 2
    assert packed.shape == (P * (P + 1) // 2,)
    assert np.array_equal(unpack_eri_s8(packed, 9), g)
    save_problem(tmp_path / "s8.npz", mol)
    save_problem(tmp_path / "dense.npz", mol, eri_packing="dense")
    with np.load(tmp_path / "s8.npz") as d:
        assert d["int2e_ao"].ndim == 1
    for name in ("s8.npz", "dense.npz"):
        m2, _ = load_problem(tmp_path / name)
        assert np.array_equal(m2.int2e_ao, g)
    bad = g.copy()
    bad[0, 1, 2, 3] += 1.0                                   # breaks the symmetry
    with pytest.raises(ValueError):
        pack_eri_s8(bad)

    class Asym:
        int1e_ao, int2e_ao, overlap, oao_coeff, nuc, nelectron = mol.int1e_ao, bad, mol.overlap, mol.oao_coeff, 0.0, 10
    save_problem(tmp_path / "auto.npz", Asym())              # auto: falls back to the dense tensor
    m3, _ = load_problem(tmp_path / "auto.npz")
    assert np.array_equal(m3.int2e_ao, bad)
    with pytest.raises(ValueError):
        save_problem(tmp_path / "x.npz", Asym(), eri_packing="s8")


def test_trajectory_checkpoint(tmp_path):
    import torch
    Cs = [torch.eye(4, dtype=torch.float64) * (i + 1) for i in range(3)]
    thetas = [torch.full((2, 1), 0.1 * i, dtype=torch.float64) for i in range(3)]
    save_trajectory(tmp_path / "t.npz", Cs, theta=thetas, energies=[-1.0, -1.5, -1.6], geometry=np.arange(3.0))
    t = load_trajectory(tmp_path / "t.npz")
    assert t["oao_mo_coeff"].shape == (3, 4, 4) and t["theta"].shape == (3, 2)
    assert np.allclose(t["energies"], [-1.0, -1.5, -1.6]) and np.array_equal(t["geometry"], np.arange(3.0))
    assert np.array_equal(t["oao_mo_coeff"][2], 3 * np.eye(4))
    with pytest.raises(ValueError):
        save_trajectory(tmp_path / "bad.npz", Cs, theta=thetas[:2])
    with pytest.raises(ValueError):
        load_trajectory(tmp_path / "missing.npz") if (tmp_path / "missing.npz").exists() else load_problem_as_trajectory(tmp_path)


def load_problem_as_trajectory(tmp_path):
    mol = SyntheticMol(4, 4, seed=0)
    save_problem(tmp_path / "p.npz", mol)
    return load_trajectory(tmp_path / "p.npz")


@pytest.mark.parametrize("nao", [8, 9])
def test_packed_loading_is_the_device_layout(tmp_path, nao):
    """eri="packed": the s8 file becomes g8[RS][PQ] over the orbitals padded to even -- the layout
    oo_pack_eri_8fold_f64 produces on the device -- without the dense tensor ever being formed."""
    mol = SyntheticMol(nao, 10, seed=4)
    g = np.asarray(mol.int2e_ao)
    save_problem(tmp_path / "s8.npz", mol)
    save_problem(tmp_path / "dense.npz", mol, eri_packing="dense")
    m2, _ = load_problem(tmp_path / "s8.npz", eri="packed")
    assert m2.int2e_ao is None
    ld = nao + (nao & 1)
    gp = np.zeros((ld,) * 4)
    gp[:nao, :nao, :nao, :nao] = g
    r, s = np.tril_indices(ld)
    expect = gp[r, s][:, r, s]
    rows = ld * (ld + 1) // 2
    assert m2.int2e_packed8.shape == (rows, rows + (rows & 1))
    assert np.array_equal(m2.int2e_packed8[:, :rows], expect) and not m2.int2e_packed8[:, rows:].any()
    with pytest.raises(ValueError):
        load_problem(tmp_path / "dense.npz", eri="packed")
    with pytest.raises(ValueError):
        load_problem(tmp_path / "s8.npz", eri="sparse")
