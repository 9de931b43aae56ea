"""Device time of kappa -> U = expm(-K) (`oo_kappa_rotation_f64`) on its two routes above 64 orbitals: the cooperative
single-launch chain and one launch per product (OO_OPT_EXPM_MULTI_LAUNCH=1, in a child process: the switch is read
once).  Prints one JSON line; `python tools/rotation_bench.py > profiles/r02_rotation_routes.json`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure():
    import numpy as np
    import torch
    from auto_oo_b200 import _lib
    from auto_oo_b200.engine import tril_pair_table
    lib = _lib.load()
    out = {}
    for N, B in ((114, 1), (114, 64), (256, 1), (256, 4)):
        ld, nk = N + (N & 1), N * (N - 1) // 2
        pl, pr = tril_pair_table(N, np.arange(nk))
        gen = torch.Generator().manual_seed(N)
        kap = (torch.randn(B, nk, dtype=torch.float64, generator=gen) * (0.5 / np.sqrt(N))).cuda()
        pld, prd = torch.as_tensor(pl).cuda(), torch.as_tensor(pr).cuda()
        K = torch.zeros(B, N, N, dtype=torch.float64)
        K[:, pl.astype(np.int64), pr.astype(np.int64)] = kap.cpu()
        s = max(0, int(np.ceil(np.log2((K - K.transpose(1, 2)).abs().sum(1).max().item() / 0.95))))
        nbytes = lib.oo_workspace_bytes(_lib.OO_WS_ROTATION, N, ld, 0, B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        U = torch.empty(B, ld, ld, dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def call():
            rc = lib.oo_kappa_rotation_f64(kap.data_ptr(), pld.data_ptr(), prd.data_ptr(), nk, N, ld, B, s,
                                           U.data_ptr(), ws.data_ptr(), nbytes, st)
            assert rc == 0, rc
        for _ in range(20):
            call()
        torch.cuda.synchronize()
        reps = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        import hashlib
        out[f"n{N}_batch{B}"] = {"squarings": s, "us_per_call": 1e3 * e0.elapsed_time(e1) / reps,
                                 "sha1_of_U": hashlib.sha1(U.cpu().numpy().tobytes()).hexdigest()}
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        print(json.dumps(measure()))
    else:
        res = {"single_launch_chain": measure()}
        env = dict(os.environ, OO_OPT_EXPM_MULTI_LAUNCH="1")
        child = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True,
                               text=True, check=True)
        res["one_launch_per_product"] = json.loads(child.stdout.strip().splitlines()[-1])
        print(json.dumps(res))
