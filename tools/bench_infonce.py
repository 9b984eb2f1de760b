#!/usr/bin/env python
"""InfoNCE forward + backward (utils/loss_functions.py:147-153) at the OpenESS sizes: own kernels vs the torch / cuBLAS formulation."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openess_b200 import _lib, losses  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    for M in (800, 1600, 3200, 6400):
        k = torch.nn.functional.normalize(torch.randn(M, 256, generator=g), dim=1).to(dev)
        q = torch.nn.functional.normalize(0.7 * k.cpu() + 0.5 * torch.randn(M, 256, generator=g), dim=1).to(dev)

        def ref():
            kk, qq = k.detach().requires_grad_(True), q.detach().requires_grad_(True)
            torch.nn.functional.cross_entropy(kk @ qq.T / 0.07, torch.arange(M, device=dev)).backward()
        with _lib.profile() as p:
            losses._infonce_raw(k, q, 0.07, True)
            torch.cuda.synchronize()
        ms = timeit(lambda: losses._infonce_raw(k, q, 0.07, True))
        torch.backends.cuda.matmul.allow_tf32 = False
        ms_ref = timeit(ref)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_ref_tf32 = timeit(ref)
        torch.backends.cuda.matmul.allow_tf32 = False
        print(json.dumps({"op": "infonce_fwd+bwd", "M": M, "D": 256, "ms_own": round(ms, 4), "ms_torch_fp32": round(ms_ref, 4),
                          "ms_torch_tf32": round(ms_ref_tf32, 4), "kernels": {n: round(v[1], 4) for n, v in p.kernels.items()}}), flush=True)


if __name__ == "__main__":
    main()
