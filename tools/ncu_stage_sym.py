"""One class-path evaluation (E+G+H) at a config shape for ncu.  Only the SECOND evaluation (warm instruction caches)
lies between cudaProfilerStart / cudaProfilerStop, so with `--profile-from-start off` exactly its launches are taken:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/ncu_stage_sym.py [workload] [auto|off] [batch]
    ncu --profile-from-start off --set full --clock-control none -o report python tools/ncu_stage_sym.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy                                                          # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic_n256_cas1212"
sym = sys.argv[2] if len(sys.argv) > 2 else "auto"
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev, eri_symmetry=sym)
eng = oo.engine
mol._int2e = mol._B = None
oo.int2e_ao = None
eng.drop_full_eri()
torch.cuda.empty_cache()
one, two = random_rdms(ncas, nelecas, seed=5, device=dev)
kap = random_kappa(oo.n_kappa, seed=3, device=dev, batch=2 * nb)
H = torch.empty(nb, eng.nk, eng.nk, dtype=torch.float64, device=dev)
Coao = eng.to_padded(oo.oao_mo_coeff, 2)
E, G, _ = eng.evaluate(Coao, one, two, kappa=kap[:nb], H_out=H)
torch.cuda.synchronize()
torch.cuda.profiler.start()
E, G, _ = eng.evaluate(Coao, one, two, kappa=kap[nb:], H_out=H)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("E", E.tolist())
