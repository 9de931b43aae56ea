#!/usr/bin/env python
"""Benchmark of the orbital-optimization hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One *evaluation* (SURVEY section 8d): given (h, g, S^-1/2, E_nuc, C_oao) resident and a rotation
kappa, gamma, Gamma -> E, the packed orbital gradient (n_kappa) and the Hessian (n_kappa^2) at
C' = S^-1/2 C_oao expm(-K(kappa)).  One *step* = one batch of ``--batch-per-gpu`` evaluations on
every rank (different kappa per evaluation; ranks hold replicas of the integrals and disjoint
kappa sets, no data-path collective => weak scaling).  Default workload: BASELINE config 5,
synthetic N=256 AO basis, CAS(12,12), random symmetric RDMs.

Prints ONE JSON line (rank 0).  ``value`` = evaluations/s with all inputs resident in HBM;
``e2e`` = the same through ``OO_energy.energy_gradient_hessian`` with host tensors (pinned H2D of
kappa/gamma/Gamma, D2H of E, G and the full Hessian inside the timed region); ``roofline`` is for
the dominant kernel, the TN-DGEMM quarter transform (2 N^5 flop per launch), against the FP64
DGEMM rate of cuBLAS measured in the same run (MEASURED_PEAKS.json has no FP64 entry; nominal B200
FP64 is 40 TFLOP/s); ``cpu_baseline`` times the CPU oracle (the reference's algorithm on torch CPU
ops, oracle/oo_oracle.py) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F64 = torch.float64
METRIC = "oo_energy_gradient_hessian_evals_per_sec"
UNIT = "evals/s"
CPU_SAMPLE_NAO = int(os.environ.get("OO_BENCH_CPU_SAMPLE_NAO", "96"))   # the CPU oracle runs the same CAS at this basis size (N=256 cannot run on a host);
                             # one E+G+H sample at N=96 is 10-15 s of work on 8-16 cores


# --------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_fp64_dgemm_peak(n=8192, reps=5):
    """cuBLAS DGEMM n^3, best of reps (the FP64 denominator; MEASURED_PEAKS.json has none)."""
    a = torch.randn(n, n, dtype=F64, device="cuda")
    b = torch.randn(n, n, dtype=F64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    del a, b, c
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best / 1e12


# --------------------------------------------------------------------------------------
_sample_cache = {}


def cpu_oracle_sample(shape_name, nao_sample, reps=1):
    """Time E + G + H of the CPU oracle (reference algorithm: three 4-index transforms and the dense
    N^6 Y-matrix per evaluation) at ``nao_sample`` orbitals with the workload's CAS, and scale to
    the workload's basis size by the reference's flop count 3*8N^5 + 6N^6."""
    from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa
    from oracle import oo_oracle as orc
    nao, nelec, ncas, nelecas = CONFIG_SHAPES[shape_name]
    ns = min(nao, nao_sample)
    nelec_s = min(nelec, 2 * (ns - ncas) + nelecas)
    nelec_s -= (nelec_s - nelecas) % 2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if (shape_name, ns) not in _sample_cache:                # input generation stays outside the timing
        mol = SyntheticMol(ns, nelec_s, seed=5)
        one, two = random_rdms(ncas, nelecas, seed=5)
        prob = orc.OracleProblem(mol.int1e_ao, mol.int2e_ao, mol.oao_coeff, mol.random_oao_mo_coeff, mol.nuc,
                                 nelec_s, ncas, nelecas, False)
        _sample_cache[(shape_name, ns)] = (prob, one, two, random_kappa(prob.n_kappa, seed=5))
    prob, one, two, kappa = _sample_cache[(shape_name, ns)]
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        prob.evaluate(one, two, kappa)
        best = min(best, time.perf_counter() - t0)
    flops = lambda n: 24.0 * n ** 5 + 6.0 * n ** 6
    scale = flops(nao) / flops(ns)
    return {"seconds_sample": best, "nao_sample": ns, "scale": scale, "cores": cores,
            "evals_per_s": 1.0 / (best * scale)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is pure
    Python over pennylane/torch and cannot be installed here -- see DESIGN.md) on the host cores."""
    if rank != 0:
        return
    from auto_oo_b200.synthetic import CONFIG_SHAPES
    nao, nelec, ncas, nelecas = CONFIG_SHAPES[args.workload]
    # keep the whole run to a few minutes whatever --steps / --warmup the driver passes
    nao_s = CPU_SAMPLE_NAO if args.steps + args.warmup <= 8 else min(CPU_SAMPLE_NAO, 80 if args.steps + args.warmup <= 24 else 64)
    for _ in range(max(1, args.warmup)):
        cpu_oracle_sample(args.workload, nao_s)
    wall, res = 0.0, None
    for _ in range(args.steps):
        res = cpu_oracle_sample(args.workload, nao_s)
        wall += res["seconds_sample"]
    value = args.steps / (wall * res["scale"])
    sample = (f"oracle E+G+H (3 four-index transforms + dense N^6 Y-matrix, reference algorithm) at "
              f"N={res['nao_sample']} with the workload's CAS({nelecas},{ncas}); scaled to N={nao} by the "
              f"reference flop count 24N^5+6N^6 (x{res['scale']:.0f})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * res["scale"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nao": nao, "cas": [nelecas, ncas]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic_n256_cas1212")
    ap.add_argument("--batch-per-gpu", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: W >= 3
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from auto_oo_b200 import OO_energy, _lib
    from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"               # the version banner goes to stdout, next to the JSON line
        from auto_oo_b200.distributed import bind_to_gpu_numa_node
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local]) if visible and visible.split(",")[local].isdigit() else local
        numa_cpus = bind_to_gpu_numa_node(phys)                # pinned staging buffers on the GPU's own node
        # NCCL writes its version banner to fd 1 when the communicator is created: keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = _lib.load()

    nao, nelec, ncas, nelecas = CONFIG_SHAPES[args.workload]
    B = args.batch_per_gpu

    # ---- synthetic inputs, generated on the device from a fixed seed (every rank: same integrals)
    mol = SyntheticMol(nao, nelec, seed=5, device=dev)
    oo = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev)
    mol._int2e = None                                      # the engine holds the only N^4 copy now
    mol._B = None
    torch.cuda.empty_cache()
    eng = oo.engine
    one, two = random_rdms(ncas, nelecas, seed=5, device=dev)
    Coao = eng.to_padded(oo.oao_mo_coeff, 2)
    nk = oo.n_kappa
    total_steps = args.warmup + args.steps
    # distinct rotations for every rank / step / batch slot
    kappas = random_kappa(nk, seed=1000 + rank, device=dev, batch=total_steps * B).reshape(total_steps, B, nk)
    squarings = eng.squarings_for(kappas.reshape(-1, nk))   # host decision made once, outside the timed region
    H_out = torch.empty(B, nk, nk, dtype=F64, device=dev)

    peak_tf = measure_fp64_dgemm_peak() if rank == 0 else 0.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- full four-index transform arm (the reference's formulation; roofline of the quarter
    #      transform kernel is measured here) ---------------------------------------------
    events = []
    for s in range(args.warmup):
        eng.evaluate(Coao, one, two, kappa=kappas[s], squarings=squarings, H_out=H_out, path="full")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.warmup, total_steps):
        E_full, G_full, H = eng.evaluate(Coao, one, two, kappa=kappas[s], squarings=squarings, H_out=H_out,
                                         transform_events=events, path="full")
    e1.record()
    barrier()
    t_full_local = e0.elapsed_time(e1) * 1e-3
    t_full = torch.tensor([t_full_local], dtype=F64, device=dev)
    if world > 1:
        dist.all_reduce(t_full, op=dist.ReduceOp.MAX)
    t_full = t_full.item()
    t_transform = sum(a.elapsed_time(b) for a, b in events) * 1e-3
    n_transforms = len(events)
    E_full, G_full = E_full.clone(), G_full.clone()
    h_diag_full = H.diagonal(dim1=1, dim2=2).sum().item()
    oo.int2e_ao = None
    eng.drop_full_eri()                                   # keep only the pair-transposed ERI copy
    torch.cuda.empty_cache()

    # ---- device-resident arm: partial (J/K class) transform ------------------------------
    for s in range(args.warmup):
        eng.evaluate(Coao, one, two, kappa=kappas[s], squarings=squarings, H_out=H_out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.oo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cls_events = []
    e0.record()
    for s in range(args.warmup, total_steps):
        E, G, H = eng.evaluate(Coao, one, two, kappa=kappas[s], squarings=squarings, H_out=H_out,
                               transform_events=cls_events)
    e1.record()
    barrier()
    launches = lib.oo_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_local = e0.elapsed_time(e1) * 1e-3
    t_dev = torch.tensor([t_local], dtype=F64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = t_dev.item()
    checksum = float(E.sum().item() + G.abs().sum().item() + H.diagonal(dim1=1, dim2=2).sum().item())
    # both formulations must agree on the last step's results
    assert (E - E_full).abs().max().item() < 1e-9 and (G - G_full).abs().max().item() < 1e-8
    assert abs(H.diagonal(dim1=1, dim2=2).sum().item() - h_diag_full) < 1e-6

    # ---- end-to-end arm: public API, host tensors in and out ------------------------------
    e2e = None
    if not args.no_e2e:
        kap_host = kappas.cpu()
        one_h, two_h = one.cpu(), two.cpu()
        for s in range(2):
            oo.energy_gradient_hessian(kap_host[s], one_h, two_h)
        barrier()
        e0.record()
        for s in range(args.warmup, total_steps):
            Eh, Gh, Hh = oo.energy_gradient_hessian(kap_host[s], one_h, two_h)
        e1.record()
        barrier()
        t_e2e = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=F64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        h2d = 8 * (B * nk + one_h.numel() + two_h.numel())
        d2h = 8 * (B + B * nk + B * nk * nk)
        e2e = {"value": world * B * args.steps / t_e2e.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h}
        assert abs(Eh[-1].item() - E[-1].item()) < 1e-9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_launch_dgemm = 4 * n_transforms
    flop_per_launch = 2.0 * nao ** 5                       # one quarter transform, algorithmic
    achieved = flop_per_launch * n_launch_dgemm / t_transform / 1e12
    roofline = {"bound": "tensor", "kernel": "dgemm_tn_kernel (quarter transform, FP64 DMMA)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": None, "flop_per_launch": flop_per_launch, "launch_ms": t_transform / n_launch_dgemm * 1e3,
                "peak_source": "cuBLAS FP64 DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 "
                               "entry); nominal B200 FP64 40 TFLOP/s => frac_of_nominal below",
                "frac_of_nominal_40tf": achieved / 40.0,
                "share_of_step": t_transform / t_full_local,
                "measured_on": "full four-index transform arm (config.full_transform_arm)"}
    prof = os.path.join(ROOT, "profiles", "dgemm_tn_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            roofline["traffic"] = json.load(f).get(args.workload)

    # the headline arm's own transform (the GEMM launches of the J/K-class transform of every evaluation)
    t_cls = sum(a.elapsed_time(b) for a, b in cls_events) * 1e-3
    ld, nIp = eng.ld, eng.nIp
    if eng.eri_is_symmetric():
        ldp, npIp = int(lib.oo_pair_ld(ld)), int(lib.oo_pair_ld(nIp))
        cls_flop = (2.0 * ld * ldp * nIp * ld + 2.0 * ldp * nIp * nIp * ld + 2.0 * ld * ld * nIp * nIp * ld
                    + 8.0 * ld ** 3 * npIp)
    else:
        cls_flop = 2.0 * ld ** 4 * nIp + 12.0 * ld ** 3 * nIp * nIp
    n_cls = B * args.steps
    roofline_class = {"bound": "tensor", "kernel": "dgemm_tn_kernel x9 (J/K-class transform of the headline arm)",
                      "achieved": cls_flop * n_cls / t_cls / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                      "frac": cls_flop * n_cls / t_cls / 1e12 / peak_tf, "flop_per_evaluation": cls_flop,
                      "ms_per_evaluation": t_cls / n_cls * 1e3, "share_of_step": t_cls / t_local,
                      "note": "algorithmic flop of the class transform (class index not padded to the 48-wide tile; "
                              "includes its three HBM-bound pack/expand passes in the time)"}

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_oracle_sample(args.workload, CPU_SAMPLE_NAO)
        cpu = {"value": r["evals_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle E+G+H at N={r['nao_sample']} (same CAS) took {r['seconds_sample']:.2f} s; scaled to "
                         f"N={nao} by the reference flop count 24N^5+6N^6 (x{r['scale']:.0f}); the reference "
                         f"algorithm cannot run at N={nao} (12 N^4 tensors = 400 GB)"}

    line = {
        "metric": METRIC, "value": world * B * args.steps / t_dev, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nao": nao, "cas": [nelecas, ncas], "n_kappa": nk,
                   "evals_per_step_per_gpu": B,
                   "numa_bound_cpus": None if numa_cpus is None else len(numa_cpus),
                   "transform": ("symmetric J/K-class transform (packed AO pairs: N^4 nI + ~7 N^3 nI^2 flop), same E/G/H"
                                 if eng.eri_is_symmetric() else
                                 "general J/K-class transform (2N^4 nI + 12 N^3 nI^2 flop), same E/G/H"),
                   "eri_symmetry_defect": eng.eri_defect,
                   "full_transform_arm": {"value": world * B * args.steps / t_full, "unit": UNIT,
                                          "transform": "full four-index (8 N^5 flop), as the reference"},
                   "l2": "inputs larger than L2 (N^4 tensors of %.1f GB)" % (nao ** 4 * 8 / 1e9)
                   if nao ** 4 * 8 > 126e6 else "inputs fit L2; distinct kappa every evaluation"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "roofline_headline_arm": roofline_class,
        "cpu_baseline": cpu, "transform_tflops": 8.0 * nao ** 5 * n_transforms / t_transform / 1e12,
        "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
