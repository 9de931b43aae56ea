"""BASELINE config 3 shape (N2 cc-pVDZ: 28 AOs, CAS(6,6)): orbital-only damped Newton-Raphson with the full
analytic orbital Hessian (OO_energy.orbital_optimization, reference oo_energy.py:426-474) -- wall time per
iteration through the public API with host tensors (what the reference's drivers do), with device tensors, and
the same loop on the verbatim-reference algorithm (CPU oracle) on the host cores.
    python tools/nr_bench.py [workload] [iterations] -> gpurun_out/nr_bench_<workload>.json"""
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy                                                            # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms                   # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "n2_ccpvdz_cas66"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
mol = SyntheticMol(nao, nelec, seed=5)
one, two = random_rdms(ncas, nelecas, seed=5)
out = {"workload": wl, "nao": nao, "cas": [nelecas, ncas], "iterations": iters}


def run(oo, a, b, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        oo.orbital_optimization(a, b, max_iterations=2, conv_tol=0.0, verbose=0, **kw)       # warm-up
        oo.oao_mo_coeff = torch.as_tensor(mol.random_oao_mo_coeff).clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = oo.orbital_optimization(a, b, max_iterations=iters, conv_tol=0.0, verbose=0, **kw)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / len(e) * 1e3, e


oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff)
out["n_kappa"] = oo.n_kappa
ms, e_host = run(oo, one, two)
out["ms_per_iteration_host_tensors"] = ms
ms, e_spec = run(oo, one, two, speculate=4)
out["ms_per_iteration_host_tensors_speculative_line_search"] = ms
ood = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff)
ms, e_dev = run(ood, one.cuda(), two.cuda())
out["ms_per_iteration_device_tensors"] = ms
out["final_energy"] = e_host[-1]
assert abs(e_host[-1] - e_spec[-1]) < 1e-9 and abs(e_host[-1] - e_dev[-1]) < 1e-9

# the reference's algorithm (CPU oracle) driven by the same NewtonStep, on the host cores
from oracle import oo_oracle as orc                                                           # noqa: E402
from auto_oo_b200.utils.newton_raphson import NewtonStep                                      # noqa: E402
torch.set_num_threads(os.cpu_count() or 1)
prob = orc.OracleProblem(mol.int1e_ao, mol.int2e_ao, mol.oao_coeff, mol.random_oao_mo_coeff, mol.nuc, nelec, ncas,
                         nelecas, False)
opt = NewtonStep(verbose=0)
n_ref = min(iters, 3)
t0 = time.perf_counter()
e_ref = []
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(n_ref):
        kappa = torch.zeros(prob.n_kappa, dtype=torch.float64)
        g = prob.gradient(one, two)
        h = prob.hessian(one, two)
        kappa, _ = opt.damped_newton_step(lambda k: prob.energy(one, two, k), (kappa,), g, h)
        prob.oao_mo_coeff = prob.oao_mo_coeff @ orc.rotation_from_kappa(kappa, prob.params_idx, prob.nao)
        e_ref.append(float(prob.energy(one, two)))
out["ms_per_iteration_cpu_oracle"] = (time.perf_counter() - t0) / n_ref * 1e3
out["cpu_cores"] = os.cpu_count()
out["energy_agrees_with_oracle_after_%d_iterations" % n_ref] = abs(e_ref[-1] - e_host[n_ref - 1])
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"nr_bench_{wl}.json"), "w"), indent=1)
