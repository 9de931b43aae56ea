"""Stress / regression for the TN-DGEMM pipeline: many trials of every tile config, dual and plain,
against cuBLAS; prints failures and the achieved TFLOP/s of the wide config."""
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
total_bad = 0
for N in (16, 32, 44, 64, 128, 256):
    B = torch.randn(K, N, dtype=torch.float64, device=dev, generator=gen)
    ref = At.T @ B
    for dual in (0, 1):
        nbad = 0
        for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
            C = torch.empty(M, N, dtype=torch.float64, device=dev)
            if dual:
                C2 = torch.empty(M, N, dtype=torch.float64, device=dev)
                lib.oo_dgemm_tn_swap02_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), C2.data_ptr(), d, d, d, N, K, M, N, N, st)
            else:
                lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, M, N, N, 1, 0, 0, 0, st)
            torch.cuda.synchronize()
            nbad += int((C - ref).abs().max().item() > 1e-9)
            if dual:
                nbad += int(not torch.equal(C2.reshape(d, d, d, N), C.reshape(d, d, d, N).permute(2, 1, 0, 3)))
        total_bad += nbad
        print(f"N={N} dual={dual}: failing trials {nbad}", flush=True)
# speed of the wide config
N = 256
M = 256 ** 3
At = torch.randn(256, M, dtype=torch.float64, device=dev)
B = torch.randn(256, N, dtype=torch.float64, device=dev)
C = torch.empty(M, N, dtype=torch.float64, device=dev)
for _ in range(2):
    lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, 256, M, N, N, 1, 0, 0, 0, st)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, 256, M, N, N, 1, 0, 0, 0, st); b.record()
    torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
print(f"quarter N=256: {best:.3f} ms  {2*256**5/best/1e9:.2f} TFLOP/s; total failing trials {total_bad}")
