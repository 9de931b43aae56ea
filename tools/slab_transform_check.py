"""Slab-parallel four-index transform over NCCL: parity against the single-GPU kernel and timing.
    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/slab_transform_check.py [N]
Every rank regenerates the same full g (seeded) for the check, cuts its slab, runs both exchange
modes, and compares its i-slab of the result with the single-GPU ``oo_int2e_transform_f64``."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from auto_oo_b200.distributed import SlabTransform          # noqa: E402
from auto_oo_b200.engine import HotPathEngine               # noqa: E402

F64 = torch.float64


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev).manual_seed(99)
    g = torch.randn(n, n, n, n, dtype=F64, device=dev, generator=gen)
    Cs = [torch.randn(n, n, dtype=F64, device=dev, generator=gen) / n ** 0.5 for _ in range(4)]
    eng = HotPathEngine.for_tensors(n, device=dev)
    ref = eng.int2e_transform(*Cs, g_ao=g)[0]
    out = {"n": n, "world": world}
    modes = ("reduce_scatter", "all_to_all", "p2p") if world > 1 else ("reduce_scatter", "all_to_all")
    for mode in modes:
        try:
            st = SlabTransform(n, mode=mode)
            st(st.take_slab(g), *Cs)
        except Exception as exc:                       # symmetric memory may be unavailable on a box
            if mode != "p2p":
                raise
            out[mode] = {"unavailable": repr(exc)[:300]}
            continue
        slab = st.take_slab(g)
        res = st(slab, *Cs)
        lo, hi = st.out_range()
        err = (res - ref[lo:hi]).abs().max()
        exact = torch.equal(res, ref[lo:hi])
        if world > 1:
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
        # timing: max over ranks of device time, 3 repetitions after one warm-up
        ts = []
        for it in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st(slab, *Cs)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it:
                ts.append(t.item())
        out[mode] = {"max_abs_err": err.item(), "bit_identical_rank0": bool(exact), "ms": min(ts),
                     "tflops_aggregate": 8 * n ** 5 / (min(ts) * 1e-3) / 1e12}
        assert err.item() < 1e-9 * n, (mode, err.item())
    # single-GPU time for the same transform
    ts = []
    for it in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.int2e_transform(*Cs, g_ao=g)
        e1.record()
        torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    out["single_gpu_ms"] = min(ts)
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"slab_transform_n{n}_g{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
