"""Timeline of the host-tensor face of OO_energy.energy_gradient_hessian at N=256 (packed Hessians, two calls in
flight): wall-clock stamps of the host side and CUDA-event stamps of compute / copies, to see what is exposed."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from auto_oo_b200 import OO_energy                                                # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa   # noqa: E402

dev = torch.device("cuda", 0)
nao, nelec, ncas, nelecas = CONFIG_SHAPES["synthetic_n256_cas1212"]
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
mol._int2e = mol._B = None
oo.int2e_ao = None
oo.engine.drop_full_eri()
if len(sys.argv) > 1 and sys.argv[1] == "unpaired":          # one quarter-1 GEMM per evaluation (A/B)
    from auto_oo_b200 import _lib
    oo.engine.flags = _lib.OO_FLAG_CLASS_Q1_UNPAIRED
one, two = random_rdms(ncas, nelecas, seed=5)
B, steps = 4, 8
kap = random_kappa(oo.n_kappa, seed=1, batch=B * steps).reshape(steps, B, -1)
kw = dict(hessian_format="packed", pinned_results=True)
for s in range(2):
    oo.energy_gradient_hessian(kap[s], one, two, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
rows = []
pending = None
for s in range(steps):
    a = time.perf_counter()
    nxt = oo.energy_gradient_hessian(kap[s], one, two, wait=False, **kw)
    b = time.perf_counter()
    if pending is not None:
        pending.wait()
    c = time.perf_counter()
    pending = nxt
    rows.append({"step": s, "call_ms": (b - a) * 1e3, "wait_prev_ms": (c - b) * 1e3, "t_end_ms": (c - t0) * 1e3})
pending.wait()
torch.cuda.synchronize()
total = (time.perf_counter() - t0) * 1e3
print(json.dumps({"rows": rows, "total_ms": total, "ms_per_step": total / steps}, indent=1))
