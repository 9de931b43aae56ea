import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d, K, N = 128, 256, 44
M = d ** 3
gen = torch.Generator(device=dev).manual_seed(1)
At = torch.randn(K, M, dtype=torch.float64, device=dev, generator=gen)
B = torch.randn(K, N, dtype=torch.float64, device=dev, generator=gen)
ref = At.T @ B
# per-kblock partial references to identify WHICH k-block is wrong
for trial in range(8):
    C = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
    C2 = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
    lib.oo_dgemm_tn_swap02_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), C2.data_ptr(), d, d, d, N, K, M, N, N, st)
    torch.cuda.synchronize()
    err = (C - ref)
    bad = (err.abs().max(dim=1).values > 1e-9).nonzero().flatten()
    if not bad.numel():
        print("trial", trial, "clean"); continue
    rows = bad.tolist()
    tiles = sorted(set(r // 256 for r in rows))
    print("trial", trial, "bad rows", len(rows), "tiles", [(t, t % 148, t // 148) for t in tiles][:8])
    for t in tiles[:3]:
        rs = [r for r in rows if r // 256 == t]
        print("   tile", t, "offsets", [r % 256 for r in rs])
        r = rs[0]
        # which k explains the error? err_row = sum_k (a_wrong - a_right)_k * B[k,:]; solve least squares for delta over k
        delta = torch.linalg.lstsq(B.T, err[r].unsqueeze(1)).solution.flatten()   # 44 eqs, 256 unknowns: min-norm
        top = torch.topk(delta.abs(), 8)
        print("   row", r, "min-norm delta top k:", [(int(i), round(float(delta[i]), 3)) for i in top.indices])
        # test hypothesis: value at some k replaced by value from another row/k-block
