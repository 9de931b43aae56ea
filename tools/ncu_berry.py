"""One batched pass over 64 geometries (config 2 shape) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/ncu_berry.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy_geometries                                               # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "ch2nh_631gs_cas44"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
mols = [SyntheticMol(nao, nelec, seed=100 + g) for g in range(G)]
one, two = random_rdms(ncas, nelecas, seed=5)
batch = OO_energy_geometries(mols, ncas, nelecas, mols[0].random_oao_mo_coeff, freeze_active=True, cuda_graphs=False)
kap = random_kappa(batch.n_kappa, seed=0, batch=G).cuda()
one, two = one.cuda(), two.cuda()
for _ in range(2):
    E, Gv, H = batch.energy_gradient_hessian(kap, one, two)
torch.cuda.synchronize()
print("E0", E[0].item())
