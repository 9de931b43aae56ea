import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d, K = 128, 256
M = d ** 3
gen = torch.Generator(device=dev).manual_seed(1)
At = torch.randn(K, M, dtype=torch.float64, device=dev, generator=gen)
for N in (16, 32, 44, 48, 64, 128):
    B = torch.randn(K, N, dtype=torch.float64, device=dev, generator=gen)
    ref = At.T @ B
    for dual in (0, 1):
        nbad, detail = 0, []
        for trial in range(12):
            C = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
            if dual:
                C2 = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
                rc = lib.oo_dgemm_tn_swap02_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), C2.data_ptr(), d, d, d, N, K, M, N, N, st)
            else:
                rc = lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, M, N, N, 1, 0, 0, 0, st)
            torch.cuda.synchronize()
            err = (C - ref).abs()
            bad = (err.max(dim=1).values > 1e-9).nonzero().flatten()
            if bad.numel():
                nbad += 1
                r0 = bad[0].item()
                cols = (err[r0] > 1e-9).nonzero().flatten().tolist()
                detail.append((bad.numel(), r0, r0 % 256, len(cols), cols[:4], float(err[r0].max())))
        print(f"N={N} dual={dual}: trials with errors {nbad}/12 {detail[:4]}", flush=True)
