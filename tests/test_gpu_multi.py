"""GPU, two or more devices: ONE evaluation sharded over the ranks of a process group
(``OO_energy(..., shard="pairs")``: every rank holds a slab of pair columns of the 8-fold packed AO integrals, builds
its additive share of the class buffer, one NCCL all-reduce) reproduces the single-GPU E, gradient and Hessian to
1e-12 relative.  Skipped on a one-GPU box; ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`` runs it."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_pair_sharded_evaluation_equals_single_gpu():
    n = min(4, torch.cuda.device_count())
    n -= n % 2
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(ROOT, "tests", "_shard_worker.py")], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SHARD_WORST_REL_DIFF" in res.stdout
