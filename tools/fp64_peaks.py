"""Measure the FP64 denominators MEASURED_PEAKS.json lacks, plus first timings of the TN-DGEMM
and the four-index transform.  Run on a B200: python tools/fp64_peaks.py [--big]
Writes gpurun_out/fp64_peaks.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import _lib  # noqa: E402

F64 = torch.float64


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sum(ts) / len(ts)


def main():
    lib = _lib.load()
    out = {"gpu": torch.cuda.get_device_name(0)}
    st = torch.cuda.current_stream().cuda_stream
    # cuBLAS DGEMM peak
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=F64, device="cuda")
        b = torch.randn(n, n, dtype=F64, device="cuda")
        c = torch.empty_like(a)
        best, mean = ev_time(lambda: torch.matmul(a, b, out=c))
        out[f"cublas_dgemm_{n}_tflops_best"] = 2 * n ** 3 / best / 1e12
        out[f"cublas_dgemm_{n}_tflops_mean"] = 2 * n ** 3 / mean / 1e12
        del a, b, c
    # sustained 3 s
    n = 8192
    a = torch.randn(n, n, dtype=F64, device="cuda"); b = torch.randn(n, n, dtype=F64, device="cuda"); c = torch.empty_like(a)
    torch.cuda.synchronize(); t0 = time.time(); k = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b, out=c)
        k += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out["cublas_dgemm_8192_tflops_sustained"] = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    # cuBLAS on the quarter-transform shape (TN: [K x M]^T [K x N]) for comparison
    sizes = [114, 256] if "--big" in sys.argv else [64, 114]
    for N in sizes:
        M = N ** 3
        At = torch.randn(N, M, dtype=F64, device="cuda")
        B = torch.randn(N, N, dtype=F64, device="cuda")
        C = torch.empty(M, N, dtype=F64, device="cuda")
        best, mean = ev_time(lambda: torch.matmul(At.T, B, out=C))
        out[f"cublas_quarter_N{N}_tflops"] = 2 * N ** 4 * N / best / 1e12
        def ours():
            rc = lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, N, M, N, N, 1, 0, 0, 0, st)
            assert rc == 0, rc
        best, mean = ev_time(ours)
        out[f"oo_dgemm_tn_quarter_N{N}_tflops"] = 2 * N ** 5 / best / 1e12
        out[f"oo_dgemm_tn_quarter_N{N}_ms"] = best * 1e3
        ref = torch.matmul(At.T[:4096], B)
        out[f"oo_dgemm_tn_quarter_N{N}_maxerr"] = (C[:4096] - ref).abs().max().item()
        del At, C
        # full transform
        g = torch.randn(N, N, N, N, dtype=F64, device="cuda")
        gm = torch.empty_like(g)
        ws = torch.empty(N ** 4 * 8, dtype=torch.uint8, device="cuda")
        def tr():
            rc = lib.oo_int2e_transform_f64(g.data_ptr(), 0, B.data_ptr(), B.data_ptr(), B.data_ptr(), B.data_ptr(),
                                            0, N, N, 1, gm.data_ptr(), ws.data_ptr(), ws.numel(), st)
            assert rc == 0, rc
        best, mean = ev_time(tr, reps=3, warm=1)
        out[f"oo_int2e_transform_N{N}_ms"] = best * 1e3
        out[f"oo_int2e_transform_N{N}_tflops"] = 8 * N ** 5 / best / 1e12
        del g, gm, ws
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp64_peaks.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
