"""Quarter 1 of the symmetric class transform at N=256 with and without its second (unpacked) store:
how much of its 6.6 ms is the 5.9 GB copy T1t it writes for the K branch?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
ld, nIp = 256, 44
ldp = int(lib.oo_pair_ld(ld))
M, K = ld * ldp, ld
At = torch.randn(K, M, dtype=torch.float64, device=dev)
B = torch.randn(K, ld, dtype=torch.float64, device=dev)
C = torch.empty(M, nIp, dtype=torch.float64, device=dev)
def t(fn):
    fn(); fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
plain = t(lambda: lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), C.data_ptr(), M, nIp, K, M, ld, nIp, 1, 0, 0, 0, st))
flop = 2.0 * M * 48 * K
print(f"plain store only: {plain:.3f} ms  {flop / plain / 1e9:.1f} TFLOP/s executed (48-wide tile)")
