"""CPU, world_size 2 over gloo: the host-side orchestration of the multi-GPU paths
(auto_oo_b200/distributed.py) -- batch sharding + gather, and the slab-parallel four-index
transform in both exchange modes with the GEMM replaced by a torch matmul."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from auto_oo_b200.distributed import (PairShard, SlabTransform, pair_slab_range, shard_range, sharded_evaluations,
                                      torch_gemm_tn)


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            got = [shard_range(n, w, r) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_pair_slab_ranges_are_even_and_cover_the_pairs():
    for ldp in (2, 78, 406, 32896):
        for w in (1, 2, 3, 8):
            got = [pair_slab_range(ldp, w, r) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == ldp
            assert all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
            assert all(lo % 2 == 0 and hi % 2 == 0 for lo, hi in got)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(123)
        g = torch.randn(n, n, n, n, dtype=torch.float64, generator=gen)
        Cs = [torch.randn(n, n, dtype=torch.float64, generator=gen) for _ in range(4)]
        ref = torch.einsum('pi,qj,rk,sl,pqrs->ijkl', *Cs, g)
        for mode in ("reduce_scatter", "all_to_all"):
            st = SlabTransform(n, mode=mode, gemm=torch_gemm_tn)
            out = st(st.take_slab(g), *Cs)
            lo, hi = st.out_range()
            err = (out - ref[lo:hi]).abs().max().item()
            assert err < 1e-10, (mode, err)
            assert tuple(st.slab_shape()) == tuple(st.take_slab(g).shape)

        # pair shard: slab bookkeeping of this rank and the all-reduce that completes the class buffer
        class _Eng:
            ld = 28
        sh = PairShard()
        sh.bind(_Eng())
        assert (sh.world, sh.rank) == (world, rank) and (sh.pq_lo, sh.pq_lo + sh.pq_cnt) == pair_slab_range(406, world, rank)
        share = torch.full((3, 4), float(rank + 1), dtype=torch.float64)
        assert torch.equal(sh.all_reduce(share), torch.full((3, 4), float(sum(range(1, world + 1))), dtype=torch.float64))

        # sharded batch: every rank ends up with the full, ordered result
        kap = torch.arange(10, dtype=torch.float64).reshape(5, 2)

        def ev(k):
            return (k.sum(dim=1), k * 2.0)

        e, g2 = sharded_evaluations(ev, kap)
        assert torch.equal(e, kap.sum(dim=1)) and torch.equal(g2, kap * 2.0)
        loc = sharded_evaluations(ev, kap, gather=False)
        lo, hi = shard_range(5, world, rank)
        assert loc[0].shape[0] == hi - lo
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 6])
def test_slab_transform_and_sharded_batch_world2(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_slab_transform_single_process():
    n = 4
    gen = torch.Generator().manual_seed(7)
    g = torch.randn(n, n, n, n, dtype=torch.float64, generator=gen)
    Cs = [torch.randn(n, n, dtype=torch.float64, generator=gen) for _ in range(4)]
    ref = torch.einsum('pi,qj,rk,sl,pqrs->ijkl', *Cs, g)
    for mode in ("reduce_scatter", "all_to_all"):
        st = SlabTransform(n, mode=mode, gemm=torch_gemm_tn)
        assert (st(st.take_slab(g), *Cs) - ref).abs().max().item() < 1e-11


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa_node never raises: without NVML / a GPU it reports None and leaves the affinity alone."""
    import os
    from auto_oo_b200.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert cpus is None and after == before or set(cpus) == after
    os.sched_setaffinity(0, before)
