"""Multi-GPU sharding of the hot path on one NVLink/NVSwitch box (one process per GPU,
``torch.distributed``).

1. :func:`shard_range` / :func:`sharded_evaluations` -- independent evaluations (a kappa sweep, the
   geometries of a Berry loop): contiguous slices of the batch per rank, integrals replicated,
   no data-path collective; results are all-gathered only if the caller wants them on every rank.

2. :class:`SlabTransform` -- ONE four-index transform (reference ``oo_energy.py:21-30``) whose N^4
   tensor is sharded over the ranks, for bases that do not fit a single GPU.  Every quarter is
   the rotating TN-GEMM of ``csrc/api.cu`` (``Out[(rest), new] = sum_lead In[lead,(rest)] C[lead,new]``);
   what differs is which index is distributed and the one exchange step:

   * ``mode="reduce_scatter"`` (the decomposition named in BASELINE.json): rank ``r`` holds the slab
     ``g[p in P_r, :, :, :]`` of the LEADING AO index.  Quarter 1 gives partial sums over ``p in P_r``
     for every destination block of the new index ``i``; block ``d`` (``i in I_d``) is summed onto rank
     ``d`` by NCCL ``reduce_scatter_tensor`` calls over row chunks, each overlapped with the GEMMs of the next
     chunk (N^4 doubles leave every rank).  Quarters 2-4 are local.
   * ``mode="all_to_all"``: rank ``r`` holds ``g[:, :, (r s) in RS_r]`` (a slab of the TRAILING pair).
     Quarters 1-2 need no communication (the contracted indices p, q are complete on every rank);
     one ``all_to_all`` re-shards ``[(rs)_loc, i, j] -> [(rs), i_loc, j]`` (N^4/G doubles leave every
     rank, and every element is summed on one rank in the single-GPU order => bit-identical results),
     then quarters 3-4 are local.

   * ``mode="p2p"``: same decomposition as ``all_to_all``, but there is no separate exchange pass:
     the output pointer of the quarter-2 GEMM for destination ``d`` is rank ``d``'s receive buffer
     mapped into this process (``torch.distributed._symmetric_memory``, NVLink peer mapping), so the
     TN-GEMM's epilogue stores travel over NVLink tile by tile while the tensor pipe works on the
     next tile -- compute and the all-to-all are ONE kernel launch per destination, bracketed by two
     stream-ordered symmetric-memory barriers.  Destinations are visited in a rank-staggered order
     so that no two ranks write to the same peer at the same time.

   All end with ``g'[i in I_r, :, :, :]`` on rank ``r``.  Per-destination GEMMs make every exchanged
   chunk contiguous, so no pack/unpack pass exists; chunk ``d+1`` is computed while chunk ``d`` is in
   flight (communication on NCCL's own stream, ``async_op=True``).

The GEMM is injectable so the orchestration is testable on CPU with the ``gloo`` backend
(``tests/test_distributed_cpu.py``); the default is the sm_100a kernel ``oo_dgemm_tn_f64``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

F64 = torch.float64


# --------------------------------------------------------------------------------------
# process placement
# --------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    the pinned staging buffers it allocates next are first-touched on that node and its device<->host copies do
    not cross the socket interconnect.  With one process per GPU, eight ranks otherwise share whatever node the
    scheduler put them on (measured on 8 x B200: 87 GB/s aggregate D2H of the Hessians).  Best effort: returns the
    CPU list, or None when NVML / the cgroup does not allow it."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                                             # noqa: BLE001  (placement is an optimisation)
        return None


# --------------------------------------------------------------------------------------
# independent evaluations
# --------------------------------------------------------------------------------------
def shard_range(n_items, world_size, rank):
    """Contiguous, balanced slice ``[lo, hi)`` of ``n_items`` for ``rank`` (first ``n % world``
    ranks get one more)."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_evaluations(evaluate, kappas, group=None, gather=True):
    """Run ``evaluate(kappa_slice) -> tuple of tensors with leading batch dim`` on this rank's
    slice of ``kappas (B, n_kappa)``; with ``gather`` every rank receives the full-batch results
    (``all_gather`` of equal-size padded slices)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = kappas.shape[0]
    lo, hi = shard_range(B, world, rank)
    local = evaluate(kappas[lo:hi])
    if not gather or world == 1:
        return local
    width = -(-B // world)
    outs = []
    for t in local:
        pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: hi - lo] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        parts = []
        for r in range(world):
            a, b = shard_range(B, world, r)
            parts.append(bufs[r][: b - a])
        outs.append(torch.cat(parts, dim=0))
    return tuple(outs)


# --------------------------------------------------------------------------------------
# one evaluation sharded over the packed AO pair index
# --------------------------------------------------------------------------------------
def pair_slab_range(n_pairs_ld, world_size, rank):
    """Columns ``[lo, hi)`` of the 8-fold packed tensor ``g8[RS][PQ]`` (``n_pairs_ld`` = padded pair count, even)
    that ``rank`` holds: contiguous, balanced, both ends even (16-byte alignment of every row slice)."""
    half = int(n_pairs_ld) // 2
    lo, hi = shard_range(half, world_size, rank)
    return 2 * lo, 2 * hi


class PairShard:
    """ONE evaluation (E + gradient + Hessian) spread over the ranks of ``group`` by the packed pair index PQ of the
    8-fold packed AO integrals ``g8[RS][PQ]`` (second decomposition; for bases whose integrals exceed one
    GPU).  Rank ``r`` holds the columns :func:`pair_slab_range`; quarter 1 of the class transform -- the N^4 nI part
    of the work and all of the N^4 memory -- runs on the slab alone, every later step is linear in its result, so
    each rank ends with an additive share of the K / J rows of the class buffer and ONE ``all_reduce`` (NCCL over
    NVLink / NVSwitch, in-switch reduction where NVLS is available) completes it on every rank.  Energy, gradient
    and Hessian are then formed on every rank from the complete class buffer (replicated, a tenth of the work).

    ``slab_given=True``: the ``int2e_ao`` handed to the engine / ``OO_energy`` IS this rank's slab
    ``(ld (ld+1)/2, hi - lo)`` (pairs over the orbitals padded to even), for integrals that no single device ever
    holds; ``False``: it is the full ``(N,N,N,N)`` tensor (tests, small problems) and the slab is cut out of it."""

    def __init__(self, group=None, slab_given=False, all_reduce=None):
        self.group, self.slab_given = group, bool(slab_given)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._all_reduce = all_reduce
        self.pq_lo = self.pq_cnt = self.slab_ld = None
        self.timing = None          # a list here collects (start, end) CUDA-event pairs around every all-reduce

    def bind(self, engine):
        from . import _lib
        ldp = int(_lib.load().oo_pair_ld(engine.ld))
        lo, hi = pair_slab_range(ldp, self.world, self.rank)
        self.pq_lo, self.pq_cnt, self.slab_ld = lo, hi - lo, hi - lo
        assert self.pq_cnt > 0, "more ranks than pair columns"

    def all_reduce(self, t):
        if self._all_reduce is not None:
            return self._all_reduce(t)
        if self.world > 1:
            if self.timing is not None and t.is_cuda:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                ev[1].record()
                self.timing.append(ev)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


# --------------------------------------------------------------------------------------
# slab-parallel four-index transform
# --------------------------------------------------------------------------------------
def _cuda_gemm_tn(At, B, out):
    """out[m, n] = sum_k At[k, m] B[k, n] on the FP64 tensor-core kernel (device tensors)."""
    from . import _lib
    lib = _lib.load()
    K, M = At.shape
    N = B.shape[1]
    rc = lib.oo_dgemm_tn_f64(At.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, At.stride(0), B.stride(0),
                             out.stride(0), 1, 0, 0, 0, torch.cuda.current_stream(At.device).cuda_stream)
    _lib.check(rc, "dgemm_tn")
    return out


def torch_gemm_tn(At, B, out):
    """Same contract on torch ops -- for the CPU/gloo tests of the orchestration only."""
    torch.matmul(At.T, B, out=out)
    return out


class SlabTransform:
    """Distributed ``general_4index_transform`` over the ranks of ``group``.

    ``n`` orbitals (even, or pre-padded), ``mode`` as in the module docstring.  ``slab_shape()``
    tells the caller which slab of the AO tensor this rank must hold; ``__call__(g_slab, C0..C3)``
    returns ``g'[i in I_r, :, :, :]`` (``out_range()``)."""

    def __init__(self, n, mode="reduce_scatter", group=None, gemm=None):
        assert mode in ("reduce_scatter", "all_to_all", "p2p")
        self.n, self.mode, self.group = int(n), mode, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gemm = gemm or _cuda_gemm_tn
        self._symm = {}
        if mode in ("all_to_all", "p2p"):
            # equal splits keep all_to_all_single simple: n^2 trailing pairs and n leading indices
            assert (self.n * self.n) % self.world == 0 and self.n % self.world == 0, \
                "all_to_all mode needs world_size | n"
        else:
            assert self.n % self.world == 0, "reduce_scatter mode needs world_size | n"
        self.blk = self.n // self.world                 # |I_d| for every destination
        if gemm is None:
            # oo_dgemm_tn_f64 needs even leading dimensions (16-byte TMA strides); the per-destination operands
            # C0[:, I_d] and the exchanged chunks have leading dimension n / world
            assert self.n % 2 == 0 and self.blk % 2 == 0, \
                f"the CUDA TN-GEMM needs n and n / world_size even (n={self.n}, world={self.world}): pad n"

    # ---- which part of g_ao each rank holds
    def in_range(self):
        """reduce_scatter: range of the leading index p; all_to_all: range of the flat (r s) pair."""
        if self.mode == "reduce_scatter":
            return shard_range(self.n, self.world, self.rank)
        return shard_range(self.n * self.n, self.world, self.rank)

    def slab_shape(self):
        lo, hi = self.in_range()
        n = self.n
        return (hi - lo, n, n, n) if self.mode == "reduce_scatter" else (n, n, hi - lo)

    def take_slab(self, g_full):
        """Cut this rank's slab out of a full (n,n,n,n) tensor (tests / small cases)."""
        lo, hi = self.in_range()
        n = self.n
        if self.mode == "reduce_scatter":
            return g_full[lo:hi].contiguous()
        return g_full.reshape(n, n, n * n)[:, :, lo:hi].contiguous()

    def out_range(self):
        return shard_range(self.n, self.world, self.rank)

    # ---- the transform
    def __call__(self, g_slab, C0, C1, C2, C3):
        if self.mode == "reduce_scatter":
            t1 = self._quarter1_reduce_scatter(g_slab, C0)
            rest = (C1, C2, C3)
        elif self.mode == "p2p" and self.world > 1:
            t1 = self._quarters12_p2p(g_slab, C0, C1)
            rest = (C2, C3)
        else:
            t1 = self._quarters12_all_to_all(g_slab, C0, C1)
            rest = (C2, C3)
        # t1 is [lead, (rest...)] with the distributed index somewhere in the middle; every
        # remaining quarter contracts the leading index against a full C and appends the new one.
        cur = t1
        for C in rest:
            lead = cur.shape[0]
            m = cur.numel() // lead
            out = torch.empty(m, self.n, dtype=F64, device=cur.device)
            self.gemm(cur.reshape(lead, m), C, out)
            cur = out.reshape(self.n, -1)                 # next leading index is the slowest of `out`
        return cur.reshape(self.blk, self.n, self.n, self.n)

    RS_CHUNKS = 8            # the reduce-scatter of quarter 1 is issued in this many row chunks (overlap with compute)

    def _quarter1_reduce_scatter(self, g_slab, C0):
        """Partial sums over this rank's p-slab for EVERY destination block of i, reduce-scattered over the ranks
        in row chunks: chunk c+1 is computed while the NCCL ``reduce_scatter`` of chunk c is on the wire."""
        n, W, blk = self.n, self.world, self.blk
        lo, hi = self.in_range()
        rows = n * n * n
        At = g_slab.reshape(hi - lo, rows)                           # [p_loc, (q r s)]
        mine = torch.empty(rows, blk, dtype=F64, device=g_slab.device)
        if W == 1:
            self.gemm(At, C0, mine)
            return mine.reshape(n, n * n * blk)
        Cd = [C0[lo:hi, d * blk:(d + 1) * blk].contiguous() for d in range(W)]     # [p_loc, i in I_d]
        step = -(-rows // self.RS_CHUNKS)
        step += step & 1                                             # even row offsets keep 16-byte alignment
        bufs = [torch.empty(W * step * blk, dtype=F64, device=g_slab.device) for _ in range(2)]
        pending = []
        for c, r0 in enumerate(range(0, rows, step)):
            r1 = min(rows, r0 + step)
            if len(pending) >= 2:                                    # the buffer we are about to reuse
                pending.pop(0).wait()
            part = bufs[c % 2][: W * (r1 - r0) * blk].view(W, r1 - r0, blk)
            for d in range(W):
                self.gemm(At[:, r0:r1], Cd[d], part[d])
            pending.append(self._reduce_scatter_async(mine[r0:r1], part))
        for w in pending:
            w.wait()
        return mine.reshape(n, n * n * blk)                          # [q, (r s i_loc)]

    def _reduce_scatter_async(self, out, parts):
        """``out = sum over ranks of parts[my rank]`` (parts: [W, ...] contiguous); returns an object with
        ``wait()``.  NCCL: ``reduce_scatter_tensor``; gloo (CPU tests) has no reduce-scatter: all-reduce + slice."""
        if dist.get_backend(self.group) == "gloo":
            work = dist.all_reduce(parts, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

            class _Then:
                def wait(_self):
                    work.wait()
                    out.copy_(parts[self.rank])
            return _Then()
        return dist.reduce_scatter_tensor(out, parts, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _quarters12_all_to_all(self, g_slab, C0, C1):
        n, W, blk = self.n, self.world, self.blk
        lo, hi = self.in_range()
        nrs = hi - lo
        At = g_slab.reshape(n, n * nrs)                              # [p, (q rs_loc)]
        send = torch.empty(W, nrs, blk, n, dtype=F64, device=g_slab.device)   # chunk d: [rs_loc, i in I_d, j]
        t1 = torch.empty(n * nrs, blk, dtype=F64, device=g_slab.device)
        for d in range(W):
            self.gemm(At, C0[:, d * blk:(d + 1) * blk].contiguous(), t1)      # [(q rs_loc), i_d]
            self.gemm(t1.reshape(n, nrs * blk), C1, send[d].reshape(nrs * blk, n))   # [(rs_loc i_d), j]
        if W == 1:
            return send[0].reshape(n, n * blk * n)
        recv = torch.empty_like(send)                                 # chunk s: [rs in RS_s, i_loc, j]
        dist.all_to_all_single(recv.reshape(-1), send.reshape(-1), group=self.group)
        return recv.reshape(n, n * blk * n)                          # [r, (s i_loc j)]

    def _symmetric_recv(self, shape, device):
        """Receive buffer of this rank, peer-mapped on every rank (allocated once per shape)."""
        import torch.distributed._symmetric_memory as symm_mem
        key = tuple(shape)
        if key not in self._symm:
            buf = symm_mem.empty(shape, dtype=F64, device=device)
            group = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, group)
            self._symm[key] = (buf, hdl)
        return self._symm[key]

    def _quarters12_p2p(self, g_slab, C0, C1):
        n, W, blk = self.n, self.world, self.blk
        lo, hi = self.in_range()
        nrs = hi - lo
        shape = (W, nrs, blk, n)                                        # chunk s: [rs in RS_s, i_loc, j]
        recv, hdl = self._symmetric_recv(shape, g_slab.device)
        At = g_slab.reshape(n, n * nrs)                                 # [p, (q rs_loc)]
        t1 = torch.empty(n * nrs, blk, dtype=F64, device=g_slab.device)
        hdl.barrier()                                                   # peers are done reading their recv buffers
        for step in range(W):
            d = (self.rank + step) % W                                  # staggered: one writer per peer at a time
            self.gemm(At, C0[:, d * blk:(d + 1) * blk].contiguous(), t1)        # [(q rs_loc), i_d]
            peer = hdl.get_buffer(d, shape, F64)                        # rank d's receive buffer, mapped here
            # quarter 2 for destination d: the GEMM epilogue writes straight into peer memory
            self.gemm(t1.reshape(n, nrs * blk), C1, peer[self.rank].reshape(nrs * blk, n))
        hdl.barrier()                                                   # every rank's stores have landed
        return recv.reshape(n, n * blk * n)                             # [r, (s i_loc j)]

    def _global_rank(self, group_rank):
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)
