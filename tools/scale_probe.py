"""Beyond BASELINE config 5: one class-path evaluation at N = 320 AOs (83.9 GB of AO integrals), CAS(12,12),
no = 48 -- checked through finite differences of the energy (the full transform no longer fits one GPU) and timed.
    python tools/scale_probe.py [N] [nelec] -> gpurun_out/scale_probe_n<N>.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy                                                       # noqa: E402
from auto_oo_b200.synthetic import SyntheticMol, random_rdms, random_kappa               # noqa: E402

F64 = torch.float64
nao = int(sys.argv[1]) if len(sys.argv) > 1 else 320
nelec = int(sys.argv[2]) if len(sys.argv) > 2 else 108
ncas = nelecas = 12
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=7, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
eng = oo.engine
assert eng.eri_is_symmetric()
eng.packed_eri()
mol._int2e = mol._B = None
oo.int2e_ao = None
eng.drop_full_eri()
torch.cuda.empty_cache()
one, two = random_rdms(ncas, nelecas, seed=7, device=dev)
nk = oo.n_kappa
Coao = eng.to_padded(oo.oao_mo_coeff, 2)
H = torch.empty(1, nk, nk, dtype=F64, device=dev)
E0, G, _ = eng.evaluate(Coao, one, two, H_out=H)
gen = torch.Generator(device=dev).manual_seed(5)
d = torch.randn(nk, dtype=F64, device=dev, generator=gen)
d /= torch.linalg.vector_norm(d)
t = 1e-2
steps = torch.tensor([-2.0, -1.0, 1.0, 2.0], dtype=F64, device=dev) * t
E, _, _ = eng.evaluate(Coao, one, two, kappa=steps[:, None] * d[None, :], want_hessian=False)
em2, em1, ep1, ep2 = (x.item() for x in E)
e0 = E0.item()
d1 = (em2 - 8 * em1 + 8 * ep1 - ep2) / (12 * t)
d2 = (-em2 + 16 * em1 - 30 * e0 + 16 * ep1 - ep2) / (12 * t * t)
g_d, h_dd = torch.dot(G[0], d).item(), torch.dot(d, H[0] @ d).item()
kap = random_kappa(nk, seed=3, device=dev, batch=1)
for _ in range(2):
    eng.evaluate(Coao, one, two, kappa=kap, H_out=H)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    eng.evaluate(Coao, one, two, kappa=kap, H_out=H)
b.record()
torch.cuda.synchronize()
out = {"nao": nao, "nelec": nelec, "cas": [12, 12], "no": eng.no, "n_kappa": nk, "eri_gb": nao ** 4 * 8 / 1e9,
       "energy": e0, "gradient_fd": d1, "gradient_analytic": g_d, "curvature_fd": d2, "curvature_analytic": h_dd,
       "hessian_asymmetry": (H[0] - H[0].T).abs().max().item(), "ms_per_evaluation": a.elapsed_time(b) / 3,
       "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9}
assert abs(d1 - g_d) < 1e-6 * max(abs(g_d), abs(h_dd), 1.0), out
assert abs(d2 - h_dd) < 1e-5 * max(abs(h_dd), 1.0), out
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"scale_probe_n{nao}.json"), "w"), indent=1)
