"""Replay of the reference's MOLECULAR golden tests (test/test_oo_energy.py) without PySCF.

``oracle/gto_sto3g.py`` rebuilds the STO-3G integrals and RHF orbitals of the reference's test molecule
(formaldimine, ``get_formal_geo(140, 80)``); ``oracle/make_molecular_golden.py`` stored them together
with the golden numbers printed in the reference's test file.  The CPU tests pin the integral code and
the oracle on those numbers; the ``gpu`` tests run the same checks through the CUDA path, written the
way the reference's tests are."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_case

F64 = torch.float64
FIX = "mol_ch2nh_sto3g_cas22"


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(GOLDEN, FIX + ".npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="module")
def rebuilt():
    from oracle.gto_sto3g import GtoMol, formaldimine_geometry
    m = GtoMol(formaldimine_geometry(140, 80))
    m.run_rhf()
    return m


# ------------------------------------------------------------------------------------------ CPU
def test_h2_sto3g_rhf_known_answer():
    """Textbook value (Szabo & Ostlund; PySCF prints -1.11675930740 for R = 0.74 A)."""
    from oracle.gto_sto3g import GtoMol
    m = GtoMol([("H", (0, 0, 0)), ("H", (0, 0, 0.74))])
    m.run_rhf()
    assert abs(m.hf.e_tot - (-1.11675930740)) < 2e-10
    assert m.nao == 2 and abs(m.overlap[0, 0] - 1) < 1e-12


def test_fixture_integrals_are_what_the_committed_code_builds(gold, rebuilt):
    assert np.abs(rebuilt.int1e_ao - gold["int1e_ao"]).max() < 1e-12
    assert np.abs(rebuilt.int2e_ao - gold["int2e_ao"]).max() < 1e-12
    assert np.abs(rebuilt.overlap - gold["overlap"]).max() < 1e-12
    assert abs(rebuilt.nuc - float(gold["nuc"])) < 1e-12
    g = rebuilt.int2e_ao                                        # 8-fold symmetry
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.abs(g - g.transpose(perm)).max() < 1e-13


def test_rhf_orbitals_in_oao_basis_match_reference_golden(gold, rebuilt):
    """reference test_mo_ao_to_oao (test/test_oo_energy.py:98-102): ``mo_ao_to_mo_oao(hf.mo_coeff, S)``
    against the printed 13x13 matrix, same tolerance (rtol 1e-5, atol 1e-8 would be sign-sensitive: the
    sign of an eigenvector is the eigensolver's choice, so columns are compared up to sign)."""
    from oracle import oo_oracle as orc
    ref = gold["reftest_hf_oao_coeff"]
    assert np.allclose(orc.mo_ao_to_mo_oao(rebuilt.oao_coeff, rebuilt.overlap), np.eye(13), atol=1e-10)
    mine = np.asarray(orc.mo_ao_to_mo_oao(rebuilt.hf.mo_coeff, rebuilt.overlap))
    sg = np.sign(np.sum(mine * ref, axis=0))
    assert np.abs(mine * sg - ref).max() < 2e-5
    # the reference's golden RHF energy (the fixed point of test_orbital_optimization)
    assert abs(rebuilt.hf.e_tot - float(gold["reftest_oo_e_ref"][0])) < 5e-8


def test_oracle_energy_from_mo_coeff_golden(gold):
    """reference test_energy_from_mo_coeff (test/test_oo_energy.py:301-308), rtol 1e-5 as there."""
    from oracle import oo_oracle as orc
    C = torch.as_tensor(gold["reftest_energy_mo_coeff"])
    h, g = orc.transform_1e(gold["int1e_ao"], C), orc.transform_2e(gold["int2e_ao"], C)
    occ, act, _ = orc.active_space_idx(13, 16, 2, 2)
    c0, c1, c2 = orc.hamiltonian_coefficients(float(gold["nuc"]), h, g, occ, act)
    e = orc.energy_from_coefficients(c0, c1, c2, torch.as_tensor(gold["reftest_energy_one_rdm"]),
                                     torch.as_tensor(gold["reftest_energy_two_rdm"])).item()
    assert np.allclose(e, gold["reftest_energy_e_ref"])                       # the reference's own assertion
    assert abs(e - float(gold["reftest_energy_value"])) < 1e-10               # verbatim reference, same inputs


def test_reference_trajectory_reaches_golden_energy(gold):
    """verbatim reference on the rebuilt integrals (stored trajectory) vs its printed e_ref
    (test/test_oo_energy.py:399-406: ``math.allclose(e_ref, energy_l[-1])``)."""
    assert np.allclose(gold["reftest_oo_e_ref"], gold["reftest_oo_trajectory"][-1])
    assert abs(gold["reftest_oo_trajectory"][-1] - float(gold["reftest_oo_e_ref"][0])) < 5e-8


# ------------------------------------------------------------------------------------------ GPU
def _mol(gold):
    from oracle.ref_shim import FakeMol
    return FakeMol(gold["int1e_ao"], gold["int2e_ao"], gold["overlap"], gold["oao_coeff"], float(gold["nuc"]), 16)


@pytest.mark.gpu
def test_cuda_energy_from_mo_coeff_golden(gold):
    import auto_oo_b200
    oo = auto_oo_b200.OO_energy(_mol(gold), 2, 2, oao_mo_coeff=gold["oao_mo_coeff"], freeze_active=True)
    e = oo.energy_from_mo_coeff(torch.as_tensor(gold["reftest_energy_mo_coeff"]),
                                torch.as_tensor(gold["reftest_energy_one_rdm"]),
                                torch.as_tensor(gold["reftest_energy_two_rdm"]))
    assert np.allclose(e.item(), gold["reftest_energy_e_ref"])
    assert abs(e.item() - float(gold["reftest_energy_value"])) < 1e-10


@pytest.mark.gpu
def test_cuda_mo_ao_to_oao_golden(gold):
    import auto_oo_b200
    assert np.allclose(np.asarray(auto_oo_b200.mo_ao_to_mo_oao(gold["oao_coeff"], gold["overlap"])), np.eye(13),
                       atol=1e-10)
    mine = np.asarray(auto_oo_b200.mo_ao_to_mo_oao(gold["hf_mo_coeff"], gold["overlap"]))
    ref = gold["reftest_hf_oao_coeff"]
    sg = np.sign(np.sum(mine * ref, axis=0))
    assert np.abs(mine * sg - ref).max() < 2e-5


@pytest.mark.gpu
def test_cuda_orbital_optimization_golden(gold):
    """reference test_orbital_optimization (test/test_oo_energy.py:399-406): closed-shell RDMs, start at
    the RHF orbitals, final energy ``allclose`` to the printed e_ref -- and the trajectory of the
    verbatim reference on the same integrals."""
    import auto_oo_b200
    oo = auto_oo_b200.OO_energy(_mol(gold), 2, 2, oao_mo_coeff=gold["oao_mo_coeff"], freeze_active=False)
    with contextlib.redirect_stdout(io.StringIO()):
        traj = oo.orbital_optimization(torch.as_tensor(gold["reftest_oo_one_rdm"]),
                                       torch.as_tensor(gold["reftest_oo_two_rdm"]))
    assert np.allclose(gold["reftest_oo_e_ref"], traj[-1])
    ref = gold["reftest_oo_trajectory"]
    assert len(traj) == len(ref) and np.abs(np.asarray(traj) - ref).max() < 1e-9


@pytest.mark.gpu
def test_cuda_int_transforms_vs_direct_contraction(gold):
    """reference test_int_transforms (test/test_oo_energy.py:114-186) compares with PySCF's ``ao2mo``;
    here the independent formula is the one-shot einsum."""
    import auto_oo_b200
    C = gold["hf_mo_coeff"]
    h, g = gold["int1e_ao"], gold["int2e_ao"]
    assert np.allclose(np.asarray(auto_oo_b200.int1e_transform(h, C)), C.T @ h @ C, rtol=0, atol=1e-11)
    direct = np.einsum('pi,qj,rk,sl,pqrs->ijkl', C, C, C, C, g, optimize=True)
    assert np.allclose(np.asarray(auto_oo_b200.int2e_transform(g, C)), direct, rtol=0, atol=1e-11)
    c = load_case("mol_ch2nh_sto3g_cas44")
    Cp = c.ref["mo_coeff_rot"]
    assert np.abs(np.asarray(auto_oo_b200.int2e_transform(c.int2e_ao, Cp)) - c.ref["int2e_mo"]).max() < 1e-11
