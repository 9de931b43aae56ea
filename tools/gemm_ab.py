"""A/B timing of the quarter-transform GEMM between two builds of the library (diagnostics)."""
import ctypes as C, sys, torch
libs = sys.argv[1:]
dev = torch.device("cuda", 0)
N = 256; M = N ** 3
At = torch.randn(N, M, dtype=torch.float64, device=dev)
B = torch.randn(N, N, dtype=torch.float64, device=dev)
Cm = torch.empty(M, N, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    for path in libs:
        lib = C.CDLL(path)
        f = lib.oo_dgemm_tn_f64
        f.argtypes = [C.c_void_p] * 3 + [C.c_int64] * 6 + [C.c_int] + [C.c_int64] * 3 + [C.c_void_p]
        for _ in range(2):
            f(At.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, N, M, N, N, 1, 0, 0, 0, st)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(At.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, N, M, N, N, 1, 0, 0, 0, st); b.record()
            torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        print(f"{path}: {best:.3f} ms {2*N**5/best/1e9:.2f} TFLOP/s", flush=True)
