import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa
wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic_n256_cas1212"
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
eng = oo.engine
mol._int2e = mol._B = None
one, two = random_rdms(ncas, nelecas, seed=5, device=dev)
kap = random_kappa(oo.n_kappa, seed=3, device=dev, batch=1)
Coao = eng.to_padded(oo.oao_mo_coeff, 2)
C = eng.mo_coeff(Coao, eng.rotation(kap))
g = eng.int2e_transform(C)[0]
cls = eng.class_integrals(C[0])
nI, nIp, ld = eng.nI, eng.nIp, eng.ld
K = cls[:nIp * nIp].reshape(nIp, nIp, ld, ld)[:nI, :nI]
J = cls[nIp * nIp:2 * nIp * nIp].reshape(nIp, nIp, ld, ld)[:nI, :nI]
dj = (J - g[:, :, :nI, :nI].permute(2, 3, 0, 1)).abs()
dk = (K - g[:, :nI, :nI, :].permute(2, 1, 0, 3)).abs()
print("J diff max", dj.max().item(), "K diff max", dk.max().item(), "|g| max", g.abs().max().item())
if dj.max().item() > 1e-8:
    idx = torch.nonzero(dj > 1e-8)
    print("J bad count", idx.shape[0], "first", idx[:5].tolist(), "last", idx[-5:].tolist())
if dk.max().item() > 1e-8:
    idx = torch.nonzero(dk > 1e-8)
    print("K bad count", idx.shape[0], "first", idx[:5].tolist(), "last", idx[-5:].tolist())
del g, dj, dk
torch.cuda.empty_cache()
Ef, Gf, Hf = eng.evaluate(Coao, one, two, kappa=kap, path="full")
Ec, Gc, Hc = eng.evaluate(Coao, one, two, kappa=kap, path="class")
print("E", Ef.item(), Ec.item(), "dE", abs(Ef.item() - Ec.item()))
print("dG", (Gf - Gc).abs().max().item(), "|G|max", Gf.abs().max().item())
print("dH", (Hf - Hc).abs().max().item(), "|H|max", Hf.abs().max().item())
