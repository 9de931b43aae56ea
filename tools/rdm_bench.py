"""Device time of the RDM extraction (forward, and forward+backward) for CAS(n,n) state vectors.
    python tools/rdm_bench.py [ncas ...] -> gpurun_out/rdm_bench.json
Reports the algorithmic work of the two kernels: excitation gather (bytes written: rows x columns x 8) and
the TN-DGEMM (2 rows ncolp^2 flop); for ncas <= 5 also the CPU oracle (reference algorithm: one sparse
mat-vec per operator)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import StatevectorRDM, _lib                                # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [4, 6, 8, 10, 12]
    rows = []
    lib = _lib.load()
    for ncas in sizes:
        D = 4 ** ncas
        gen = torch.Generator(device="cuda").manual_seed(ncas)
        psi = torch.randn(D, dtype=torch.complex128, device="cuda", generator=gen)
        psi = psi / torch.linalg.vector_norm(psi)
        rdm = StatevectorRDM(ncas, chunk=1 << 18)
        ncolp = int(lib.oo_rdm_columns(ncas))
        fwd = timed(lambda: rdm.get_rdms_from_state(psi))
        row = {"ncas": ncas, "amplitudes": D, "forward_ms": fwd,
               "gemm_flop": 2.0 * 2 * D * ncolp * ncolp, "phi_bytes": 8.0 * 2 * D * ncolp,
               "gemm_tflops_if_all_time": 2.0 * 2 * D * ncolp * ncolp / fwd / 1e9,
               "phi_gbs_if_all_time": 8.0 * 2 * D * ncolp / fwd / 1e6}
        if ncas <= 10:
            p2 = psi.clone().requires_grad_(True)

            def fb():
                one, two = rdm.get_rdms_from_state(p2)
                (one.sum() + (two * two).sum()).backward()
                p2.grad = None
            row["forward_backward_ms"] = timed(fb)
        if ncas <= 5:
            from oracle import rdm_oracle as ro
            t0 = time.perf_counter()
            ro.rdms_from_state(psi.cpu().numpy(), ncas)
            row["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        # the physical case: a state of the half-filled (ncas/2 up, ncas/2 down) sector, which every
        # number-conserving ansatz of the reference produces -- the passes run on that sector's basis states only
        nq = 2 * ncas
        x = torch.arange(D, device="cuda", dtype=torch.int64)
        nup = sum(((x >> (nq - 1 - 2 * p)) & 1) for p in range(ncas))
        ndn = sum(((x >> (nq - 2 - 2 * p)) & 1) for p in range(ncas))
        keep = (nup == ncas // 2) & (ndn == ncas - ncas // 2)
        ps = torch.where(keep, psi, torch.zeros_like(psi))
        ps = ps / torch.linalg.vector_norm(ps)
        row["sector_states"] = int(keep.sum().item())
        row["sector_forward_ms"] = timed(lambda: rdm.get_rdms_from_state(ps))
        assert rdm._k.last_rows == row["sector_states"]
        p3 = ps.clone().requires_grad_(True)

        def fbs():
            one, two = rdm.get_rdms_from_state(p3)
            (one.sum() + (two * two).sum()).backward()
            p3.grad = None
        row["sector_forward_backward_ms"] = timed(fbs)
        rows.append(row)
        print(row, flush=True)
        del rdm, psi, ps, p3, x, nup, ndn, keep
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "rdm_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
