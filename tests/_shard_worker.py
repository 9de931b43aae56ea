"""Worker of tests/test_gpu_multi.py (one process per GPU under torchrun): the pair-sharded evaluation against the
single-GPU evaluation of the same problem, on every rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from auto_oo_b200 import OO_energy                                                           # noqa: E402
from auto_oo_b200.synthetic import SyntheticMol, random_rdms, random_kappa                  # noqa: E402
from oracle.ref_shim import FakeMol                                                          # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    worst = 0.0
    # (nao, nelec, ncas, nelecas, slab straight from the density-fitting factor?)
    for nao, nelec, ncas, nelecas, from_factor in [(13, 16, 2, 2, False), (28, 14, 6, 6, True), (57, 36, 4, 4, False),
                                                   (114, 42, 6, 6, True)]:
        mol = SyntheticMol(nao, nelec, seed=7)
        one, two = random_rdms(ncas, nelecas, seed=7)
        single = OO_energy(mol, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev, cuda_graphs=False)
        kappa = random_kappa(single.n_kappa, seed=7, batch=2)
        E1, G1, H1 = single.energy_gradient_hessian(kappa, one, two)
        src = mol if from_factor else FakeMol(mol.int1e_ao, mol.int2e_ao, mol.overlap, mol.oao_coeff, mol.nuc, nelec)
        shard = OO_energy(src, ncas, nelecas, oao_mo_coeff=mol.random_oao_mo_coeff, device=dev, shard="pairs")
        sh = shard.engine.pair_shard
        assert sh.world == world and sh.rank == rank and sh.slab_given == from_factor
        assert tuple(shard.engine.packed_eri().shape) == (shard.engine.ld * (shard.engine.ld + 1) // 2, sh.slab_ld)
        E2, G2, H2 = shard.energy_gradient_hessian(kappa, one, two)
        scale = max(1.0, H1.abs().max().item())
        d = max((E1 - E2).abs().max().item() / max(1.0, E1.abs().max().item()), (G1 - G2).abs().max().item() / scale,
                (H1 - H2).abs().max().item() / scale)
        worst = max(worst, d)
        # the other entry points go through the same sharded transform
        e = shard.energy_from_kappa(kappa[0], one, two).item()
        g = shard.kappa_matrix_to_vector(shard.analytic_gradient(one, two))
        assert abs(e - E1[0].item()) < 1e-10 and g.shape == G1[0].shape
        del single, shard
        torch.cuda.empty_cache()
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"SHARD_WORST_REL_DIFF {t.item():.3e}", flush=True)
    dist.destroy_process_group()
    assert t.item() < 1e-12, t.item()


if __name__ == "__main__":
    main()
