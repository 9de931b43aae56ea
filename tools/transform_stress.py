"""Stress the class transform's staged epilogues / bulk-copy pipelines: the same transform repeated many times must
give bit-identical class buffers every time (a stage or tile buffer overwritten under a pending read shows up as a
sporadic mismatch), and must agree with the direct-store variant of the same kernels to round-off.
    python tools/transform_stress.py [workload] [repetitions]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy, _lib                                                   # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_kappa              # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic_n256_cas1212"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
eng = oo.engine
mol._int2e = mol._B = None
oo.int2e_ao = None
eng.drop_full_eri()
Coao = eng.to_padded(oo.oao_mo_coeff, 2)
bad, worst = 0, 0.0
for block in range(4):
    kap = random_kappa(oo.n_kappa, seed=block, device=dev, batch=1)
    C = eng.mo_coeff(Coao, eng.rotation(kap))
    ref = eng.class_integrals(C).clone()
    eng.flags = _lib.OO_FLAG_CLASS_DIRECT_STORES
    direct = eng.class_integrals(C).clone()
    eng.flags = 0
    worst = max(worst, ((ref - direct).abs().max() / ref.abs().max()).item())
    out = torch.empty_like(ref)
    for _ in range(reps // 4):
        eng.class_integrals(C, out=out)
        if not torch.equal(out, ref):
            bad += 1
print(json.dumps({"workload": wl, "repetitions": 4 * (reps // 4), "mismatching_repetitions": bad,
                  "staged_vs_direct_max_rel_diff": worst}))
sys.exit(1 if bad or worst > 1e-12 else 0)
