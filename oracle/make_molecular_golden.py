"""TEST INFRASTRUCTURE ONLY -- molecular golden fixtures (``tests/golden/mol_*.npz``).

    python -m oracle.make_molecular_golden          (build container only)

What it does:
1. builds the STO-3G integrals of the reference's test molecule (formaldimine at the two geometries
   the reference tests use, ``test/test_oo_energy.py:30,108-111``) and of water (BASELINE config 1)
   with ``oracle/gto_sto3g.py``, plus the RHF orbitals;
2. reads the golden vectors PRINTED IN THE REFERENCE'S OWN TESTS (``test/test_oo_energy.py``: RHF
   orbitals in the OAO basis ``:27-103``; ``mo_coeff`` / RDMs / ``e_ref`` of
   ``test_energy_from_mo_coeff`` ``:240-315`` and ``test_orbital_optimization`` ``:317-412``; the STO-3G
   case of ``test_analytical_derivatives`` ``:415-473``) by parsing the test file's ``parametrize``
   decorators -- data only, no code is taken -- and stores them next to the integrals;
3. runs the VERBATIM reference (``oracle/ref_shim.py``) on those molecular inputs and stores its
   outputs in the same layout as the synthetic fixtures of ``oracle/make_golden.py``, so every
   parametrised parity test (oracle on CPU, CUDA path on the GPU) also runs on real molecules;
4. prints how well the rebuilt integrals reproduce the reference's printed numbers (this is what
   pins ``gto_sto3g.py``; the same checks are asserted by ``tests/test_molecular_golden.py``).
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from auto_oo_b200.synthetic import random_rdms, random_kappa, CIVectorCircuit     # noqa: E402
from oracle.gto_sto3g import GtoMol, formaldimine_geometry, water_geometry        # noqa: E402
from oracle.make_golden import reference_case, GOLDEN                             # noqa: E402
from oracle.ref_shim import load_reference, REF_ROOT                              # noqa: E402


# ------------------------------------------------------------------------------------------
def reference_test_parameters(test_file, function_name):
    """Evaluate the ``@pytest.mark.parametrize`` argument list of ``function_name`` in the
    reference's test file.  ``math.array`` -> numpy, ``auto_oo.get_formal_geo(a, p)`` -> ("formal", a, p)."""
    with open(test_file) as f:
        tree = ast.parse(f.read())

    class _Math:
        array = staticmethod(lambda x, **kw: np.array(x))

    class _AutoOO:
        get_formal_geo = staticmethod(lambda a, p: ("formal", a, p))

    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == function_name:
            for dec in node.decorator_list:
                if isinstance(dec, ast.Call) and ast.unparse(dec.func).endswith("parametrize"):
                    names = ast.literal_eval(dec.args[0])
                    if isinstance(names, str):
                        names = [n.strip() for n in names.split(",")]
                    values = eval(compile(ast.Expression(dec.args[1]), test_file, "eval"),
                                  {"math": _Math, "auto_oo": _AutoOO, "np": np})
                    return [dict(zip(names, v)) for v in values]
    raise KeyError(function_name)


def save(name, out):
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)


def main():
    torch.set_default_dtype(torch.float64)
    ref = load_reference()
    tfile = os.path.join(REF_ROOT, "test", "test_oo_energy.py")
    mo_ao_to_mo_oao = ref.oo_energy.mo_ao_to_mo_oao

    mols = {}

    def mol_for(a, p):
        if (a, p) not in mols:
            m = GtoMol(formaldimine_geometry(a, p))
            m.run_rhf()
            mols[(a, p)] = m
        return mols[(a, p)]

    # ---------------- formaldimine (140, 80): the reference's printed goldens ---------------
    m = mol_for(140, 80)
    C_oao_hf = mo_ao_to_mo_oao(m.hf.mo_coeff, m.overlap)                   # default orbitals, oo_energy.py:143-147
    extra = dict(hf_mo_coeff=m.hf.mo_coeff, hf_e_tot=np.asarray(m.hf.e_tot))

    p = reference_test_parameters(tfile, "test_mo_ao_to_oao")[0]
    assert p["geometry"] == ("formal", 140, 80) and p["basis"] == "sto-3g"
    ref_oao = p["hf_oao_coeff_ref"]
    sg = np.sign(np.sum(ref_oao * C_oao_hf, axis=0))
    print("RHF orbitals in the OAO basis vs test_mo_ao_to_oao (up to column sign): max |d| = %.2e"
          % np.abs(C_oao_hf * sg - ref_oao).max())
    extra["reftest_hf_oao_coeff"] = ref_oao

    p = reference_test_parameters(tfile, "test_energy_from_mo_coeff")[0]
    assert p["geometry"] == ("formal", 140, 80) and (p["ncas"], p["nelecas"], p["freeze_active"]) == (2, 2, True)
    oo = ref.oo_energy.OO_energy(m, 2, 2, oao_mo_coeff=C_oao_hf, freeze_active=True, interface='torch')
    e = oo.energy_from_mo_coeff(torch.as_tensor(p["mo_coeff"]), torch.as_tensor(p["one_rdm"]),
                                torch.as_tensor(p["two_rdm"])).item()
    print("test_energy_from_mo_coeff: E = %.10f  e_ref = %.10f  (reference tolerance rtol 1e-5)"
          % (e, float(p["e_ref"][0])))
    extra.update(reftest_energy_mo_coeff=p["mo_coeff"], reftest_energy_one_rdm=p["one_rdm"],
                 reftest_energy_two_rdm=p["two_rdm"], reftest_energy_e_ref=p["e_ref"],
                 reftest_energy_value=np.asarray(e))

    p = reference_test_parameters(tfile, "test_orbital_optimization")[0]
    assert p["geometry"] == ("formal", 140, 80) and (p["ncas"], p["nelecas"], p["freeze_active"]) == (2, 2, False)
    import contextlib, io
    oo = ref.oo_energy.OO_energy(m, 2, 2, oao_mo_coeff=C_oao_hf, freeze_active=False, interface='torch')
    with contextlib.redirect_stdout(io.StringIO()):
        traj = oo.orbital_optimization(torch.as_tensor(p["one_rdm"]), torch.as_tensor(p["two_rdm"]))
    print("test_orbital_optimization: E_final = %.12f  e_ref = %.12f  (RHF here %.12f)"
          % (traj[-1], float(p["e_ref"][0]), m.hf.e_tot))
    extra.update(reftest_oo_one_rdm=p["one_rdm"], reftest_oo_two_rdm=p["two_rdm"], reftest_oo_e_ref=p["e_ref"],
                 reftest_oo_trajectory=np.asarray(traj))

    p = reference_test_parameters(tfile, "test_analytical_derivatives")[0]
    assert p["geometry"] == ("formal", 140, 80) and p["basis"] == "sto-3g"
    extra.update(reftest_deriv_one_rdm=p["one_rdm"], reftest_deriv_two_rdm=p["two_rdm"],
                 reftest_deriv_shape=np.array([p["ncas"], p["nelecas"], int(p["freeze_active"])]))

    # fixture in the common layout: CAS(2,2) frozen-active, the reference's printed (4-digit) RDMs
    # symmetrised so that the analytic formulas' preconditions hold exactly
    one = torch.as_tensor(p["one_rdm"])
    two = torch.as_tensor(p["two_rdm"])
    one = 0.5 * (one + one.T)
    two = 0.25 * (two + two.permute(2, 3, 0, 1) + two.permute(1, 0, 3, 2) + two.permute(3, 2, 1, 0))
    out, d, nk = reference_case(ref, m, m.nelectron, 2, 2, True, C_oao_hf,
                                lambda n: random_kappa(n, seed=11, scale=0.05), one, two, True, seed=11,
                                trajectory=True)
    out.update(extra)
    save("mol_ch2nh_sto3g_cas22", out)
    print(f"mol_ch2nh_sto3g_cas22   nk={nk} E={float(out['E']):+.10f} oracle-vs-reference {max(d):.1e}")

    # ---------------- formaldimine (120, 125), CAS(4,4), free active rotations ----------------
    m2 = mol_for(120, 125)
    C2 = mo_ao_to_mo_oao(m2.hf.mo_coeff, m2.overlap)
    circ = CIVectorCircuit(4, 4, n_theta=3, seed=3)
    one, two = circ.get_rdms(torch.tensor([0.4, -0.2, 0.1]))
    out, d, nk = reference_case(ref, m2, m2.nelectron, 4, 4, False, C2,
                                lambda n: random_kappa(n, seed=12, scale=0.1), one.detach(), two.detach(), True,
                                seed=12, trajectory=True)
    out.update(hf_mo_coeff=m2.hf.mo_coeff, hf_e_tot=np.asarray(m2.hf.e_tot))
    save("mol_ch2nh_sto3g_cas44", out)
    print(f"mol_ch2nh_sto3g_cas44   nk={nk} E={float(out['E']):+.10f} oracle-vs-reference {max(d):.1e}  "
          f"RHF {m2.hf.e_tot:.10f}")

    # ---------------- water, STO-3G, CAS(4,4): BASELINE config 1 (nv = 0) ----------------------
    w = GtoMol(water_geometry())
    w.run_rhf()
    Cw = mo_ao_to_mo_oao(w.hf.mo_coeff, w.overlap)
    circ = CIVectorCircuit(4, 4, n_theta=3, seed=4)
    one, two = circ.get_rdms(torch.tensor([0.3, 0.15, -0.25]))
    out, d, nk = reference_case(ref, w, w.nelectron, 4, 4, False, Cw,
                                lambda n: random_kappa(n, seed=13, scale=0.1), one.detach(), two.detach(), True,
                                seed=13, trajectory=True)
    out.update(hf_mo_coeff=w.hf.mo_coeff, hf_e_tot=np.asarray(w.hf.e_tot))
    save("mol_h2o_sto3g_cas44", out)
    print(f"mol_h2o_sto3g_cas44     nk={nk} E={float(out['E']):+.10f} oracle-vs-reference {max(d):.1e}  "
          f"RHF {w.hf.e_tot:.10f}")


if __name__ == "__main__":
    main()
