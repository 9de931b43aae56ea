"""Per-stage device times (CUDA events) of one class-path evaluation at a config shape.
    python tools/stage_times.py [workload] [auto|off] -> gpurun_out/stage_times_<workload>[_general].json
(second argument: eri_symmetry; "off" times the general class transform quarter by quarter)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy, _lib                                     # noqa: E402
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa  # noqa: E402

F64 = torch.float64


class Timer:
    def __init__(self):
        self.rows = []

    def __call__(self, name, fn, flop=None, bytes_=None):
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        row = {"stage": name, "ms": best}
        if flop:
            row["tflops"] = flop / best / 1e9
        if bytes_:
            row["gbs"] = bytes_ / best / 1e6
        self.rows.append(row)
        print(row, flush=True)
        return out


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic_n256_cas1212"
    sym = sys.argv[2] if len(sys.argv) > 2 else "auto"
    nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
    dev = torch.device("cuda", 0)
    mol = SyntheticMol(nao, nelec, seed=5, device=dev)
    oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev, eri_symmetry=sym)
    eng, lib = oo.engine, _lib.load()
    mol._int2e = mol._B = None
    oo.int2e_ao = None
    eng.drop_full_eri()
    torch.cuda.empty_cache()
    one, two = random_rdms(ncas, nelecas, seed=5, device=dev)
    kap = random_kappa(oo.n_kappa, seed=3, device=dev, batch=1)
    ld, nIp, N, nk = eng.ld, eng.nIp, eng.N, eng.nk
    st = lambda: torch.cuda.current_stream().cuda_stream
    t = Timer()
    U = t("rotation(expm)", lambda: eng.rotation(kap))
    C = t("mo_coeff", lambda: eng.mo_coeff(eng.to_padded(oo.oao_mo_coeff, 2), U))[0]
    ld2, ld3, nI2 = ld * ld, ld ** 3, nIp * nIp
    cls = torch.empty(1, 2 * nI2 + 1, ld, ld, dtype=F64, device=dev)
    if eng.eri_is_symmetric():
        ldp, npIp = int(lib.oo_pair_ld(ld)), int(lib.oo_pair_ld(nIp))
        flop = 2.0 * ld * ldp * nIp * ld + 2.0 * ldp * nI2 * ld + 2.0 * ld2 * nI2 * ld + 8.0 * ld3 * npIp
        t("class_transform_sym (all; per-kernel split: ncu launch list)", lambda: eng.class_integrals(C, out=cls),
          flop=flop)
    else:
        gp = eng.pair_transposed_eri()
        # the GEMMs of the general class transform, individually
        T1 = torch.empty(ld3 * nIp, dtype=F64, device=dev)
        X = torch.empty(ld2 * nI2, dtype=F64, device=dev)
        Xp = torch.empty_like(X)

        def gemm(a, out, M, Nc):
            rc = lib.oo_dgemm_tn_f64(a.data_ptr(), C.data_ptr(), out.data_ptr(), M, Nc, ld, M, ld, Nc, 1, 0, 0, 0,
                                     st())
            assert rc == 0

        t("Q1  [ld^3 x nIp x ld]", lambda: gemm(gp, T1, ld3, nIp), flop=2.0 * ld3 * nIp * ld,
          bytes_=8.0 * (ld ** 4 + ld3 * nIp))
        t("Q2  [ld^2 nIp x nIp x ld]", lambda: gemm(T1, X, ld2 * nIp, nIp), flop=2.0 * ld2 * nIp * nIp * ld)
        t("Q3  [ld nIp^2 x ld x ld]", lambda: gemm(X, Xp, ld * nI2, ld), flop=2.0 * ld * nI2 * ld * ld)
        t("Q4  [nIp^2 ld x ld x ld]", lambda: gemm(Xp, cls[0, nI2:], nI2 * ld, ld), flop=2.0 * ld * nI2 * ld * ld)
        del T1, X, Xp
        t("class_transform (all)", lambda: eng.class_integrals(C, out=cls),
          flop=2.0 * ld ** 4 * nIp + 12.0 * ld3 * nI2)
    c = t("active_hamiltonian", lambda: eng.class_active_hamiltonian(cls))
    t("energy", lambda: eng.energy(*c, one, two))
    FI, FA, F, _, gv = t("fock+gradient", lambda: eng.class_fock_gradient(cls, one, two, want_matrix=False))
    H = torch.empty(nk, nk, dtype=F64, device=dev)
    t("hessian (At + GEMM + assemble)", lambda: eng.class_hessian(cls, F, one, two, out=H[None]),
      flop=2.0 * nI2 * ld2 * (2 * nI2 + 1))
    t("evaluate (E+G+H)", lambda: eng.evaluate(eng.to_padded(oo.oao_mo_coeff, 2), one, two, kappa=kap, H_out=H[None]))
    Coao = eng.to_padded(oo.oao_mo_coeff, 2)
    t("evaluate_graphed (E+G+H, one CUDA-graph launch)", lambda: eng.evaluate_graphed(Coao, one, two, kappa=kap,
                                                                                       clone=False))
    for B in (8, 64):
        if nao > 128:
            break
        kb = random_kappa(oo.n_kappa, seed=4, device=dev, batch=B)
        r = t(f"evaluate batch of {B} (direct)", lambda: eng.evaluate(Coao, one, two, kappa=kb))
        t.rows[-1]["evals_per_s"] = B / t.rows[-1]["ms"] * 1e3
        r = t(f"evaluate batch of {B} (graph)", lambda: eng.evaluate_graphed(Coao, one, two, kappa=kb, clone=False))
        t.rows[-1]["evals_per_s"] = B / t.rows[-1]["ms"] * 1e3
        del r
    # the public API with host tensors (what NewtonStep / OO_pqc call), wall clock per call
    import time
    kh, oh, th = kap.cpu(), one.cpu(), two.cpu()
    for graphs in (False, True):
        oo.cuda_graphs = graphs
        for _ in range(3):
            oo.energy_gradient_hessian(kh, oh, th)
        t0 = time.perf_counter()
        for _ in range(10):
            oo.energy_gradient_hessian(kh, oh, th)
        row = {"stage": f"OO_energy.energy_gradient_hessian host->host, cuda_graphs={graphs} (wall)",
               "ms": (time.perf_counter() - t0) / 10 * 1e3}
        t.rows.append(row)
        print(row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "" if eng.eri_is_symmetric() else "_general"
    with open(os.path.join(ROOT, "gpurun_out", f"stage_times_{wl}{tag}.json"), "w") as f:
        json.dump(t.rows, f, indent=1)


if __name__ == "__main__":
    main()
