import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy, _lib
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa
nao = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nelec, ncas, nelecas = 76, 12, 12
dev = torch.device("cuda", 0)
mol = SyntheticMol(nao, nelec, seed=5, device=dev)
oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
eng = oo.engine
lib = _lib.load()
mol._int2e = mol._B = None
kap = random_kappa(oo.n_kappa, seed=3, device=dev, batch=1)
C = eng.mo_coeff(eng.to_padded(oo.oao_mo_coeff, 2), eng.rotation(kap))[0]
nI, nIp, ld = eng.nI, eng.nIp, eng.ld
gp = eng.pair_transposed_eri()
print("transpose ok:", torch.equal(gp.reshape(ld*ld, ld*ld)[:1000], eng.g_ao.reshape(ld*ld, ld*ld).T[:1000]))
cls = eng.class_integrals(C)
ws = eng._ws["cls"].view(torch.float64)
ld3 = ld ** 3
T1 = ws[:ld3 * nIp].reshape(ld3, nIp)
T1t = ws[ld3 * nIp:2 * ld3 * nIp].reshape(ld, ld, ld, nIp)
gpm = gp.reshape(ld, ld3)
for lo in (0, 1000000, ld3 - 4096):
    ref = gpm[:, lo:lo + 4096].T @ C[:, :nIp]
    d = (T1[lo:lo + 4096] - ref).abs()
    print("Q1 rows", lo, "max diff", d.max().item(), "cols with err", (d.max(0).values > 1e-10).nonzero().flatten().tolist()[:50])
T1v = T1.reshape(ld, ld, ld, nIp)
print("T1t == swap(T1):", torch.equal(T1t, T1v.permute(2, 1, 0, 3)))
# standalone gemm N=44 vs N=64-tile (force by calling with N=nIp on a fresh buffer)
out = torch.empty(ld3, nIp, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
rc = lib.oo_dgemm_tn_f64(gp.data_ptr(), C.data_ptr(), out.data_ptr(), ld3, nIp, ld, ld3, ld, nIp, 1, 0, 0, 0, st)
print("standalone rc", rc, "equal to T1:", torch.equal(out, T1))
ref = gpm[:, :4096].T @ C[:, :nIp]
print("standalone diff", (out[:4096] - ref).abs().max().item())
