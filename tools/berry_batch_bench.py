"""BASELINE config 2 shape: formaldimine 6-31G* CAS(4,4), 64 geometries -- E + gradient + Hessian of
every geometry in one batched pass (OO_energy_geometries), sharded over ranks when launched with
torchrun.  Prints evaluations/s (device-resident and with host tensors)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200 import OO_energy_geometries
from auto_oo_b200.distributed import shard_range
from auto_oo_b200.synthetic import CONFIG_SHAPES, SyntheticMol, random_rdms, random_kappa

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "ch2nh_631gs_cas44"
G_total = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nao, nelec, ncas, nelecas = CONFIG_SHAPES[wl]
lo, hi = shard_range(G_total, world, rank)
mols = [SyntheticMol(nao, nelec, seed=100 + g) for g in range(lo, hi)]
one, two = random_rdms(ncas, nelecas, seed=5)
batch = OO_energy_geometries(mols, ncas, nelecas, mols[0].random_oao_mo_coeff, freeze_active=True)
kap = random_kappa(batch.n_kappa, seed=rank, batch=hi - lo)
kap_d, one_d, two_d = kap.cuda(), one.cuda(), two.cuda()
for _ in range(3):
    batch.energy_gradient_hessian(kap_d, one_d, two_d)
torch.cuda.synchronize()
reps = 20
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    E, Gv, H = batch.energy_gradient_hessian(kap_d, one_d, two_d)
b.record(); torch.cuda.synchronize()
t_dev = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
t0 = time.perf_counter()
for _ in range(reps):
    batch.energy_gradient_hessian(kap, one, two)
t_host = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device="cuda")
if world > 1:
    dist.all_reduce(t_dev, op=dist.ReduceOp.MAX); dist.all_reduce(t_host, op=dist.ReduceOp.MAX)
if rank == 0:
    out = {"workload": wl, "geometries": G_total, "n_gpus": world, "n_kappa": batch.n_kappa,
           "ms_per_pass_device": t_dev.item(), "evals_per_s_device": G_total / t_dev.item() * 1e3,
           "ms_per_pass_host_tensors": t_host.item(), "evals_per_s_host_tensors": G_total / t_host.item() * 1e3}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"berry_batch_{wl}_g{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
