"""Device time of the symmetric class transform of a batch under the A/B flags of `oo_class_transform_sym_f64`
(staged vs direct epilogues, paired vs one-per-evaluation quarter 1), whole and quarter 1 alone.
    python tools/transform_variants.py [workload] [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy, _lib                                                    # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_kappa                # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c6h6_ccpvdz_cas66"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
eng = oo.engine
mol._int2e = mol._B = None
oo.int2e_ao = None
eng.drop_full_eri()
kap = random_kappa(oo.n_kappa, seed=3, device=dev, batch=B)
C = eng.mo_coeff(eng.to_padded(oo.oao_mo_coeff, 2), eng.rotation(kap))
variants = {"default": 0, "direct_stores": _lib.OO_FLAG_CLASS_DIRECT_STORES, "q1_unpaired": _lib.OO_FLAG_CLASS_Q1_UNPAIRED,
            "q1_unpaired+direct_stores": _lib.OO_FLAG_CLASS_Q1_UNPAIRED | _lib.OO_FLAG_CLASS_DIRECT_STORES}
out = {"workload": wl, "batch": B, "us_per_evaluation": {}}
cls = None
for name, fl in variants.items():
    row = {}
    for what, stage in (("transform", 0), ("quarter_1", _lib.OO_FLAG_CLASS_STAGE(0)), ("coulomb_q2", _lib.OO_FLAG_CLASS_STAGE(1)),
                        ("exchange_q2", _lib.OO_FLAG_CLASS_STAGE(4))):
        eng.flags = fl | stage
        for _ in range(3):
            cls = eng.class_integrals(C, out=cls)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            cls = eng.class_integrals(C, out=cls)
        e1.record()
        torch.cuda.synchronize()
        row[what] = round(1e3 * e0.elapsed_time(e1) / reps / B, 2)
    out["us_per_evaluation"][name] = row
eng.flags = 0
print(json.dumps(out))
