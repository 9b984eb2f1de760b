#!/usr/bin/env python
"""GEMM shapes of the ViT / teacher decoder through oess_gemm_tf32_ex (plain, GELU, residual) + the MaskCLIP forward.
OESS_GEMM_DEEP=1 selects the one-CTA-per-SM 4-stage variant for comparison."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import ops  # noqa: E402
from tools.bench_tc import timeit  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    for (M, N, K, act, res) in [(8968, 2304, 768, None, False), (8968, 768, 768, None, True), (8968, 3072, 768, "gelu", False),
                                (8968, 768, 3072, None, True), (140800, 256, 2048, None, False), (8960, 512, 768, None, False)]:
        a = torch.randn(M, K, device="cuda")
        b = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if res else None
        out = torch.empty(M, N, device="cuda")
        t = timeit(lambda: ops.gemm_tf32_ex(a, b, bias, residual=r, act=act, out=out), iters=30, warm=5)
        tl = timeit(lambda: torch.addmm(bias, a, b.t(), out=out), iters=30, warm=5)
        print(json.dumps({"op": "gemm_tf32_ex", "deep": os.environ.get("OESS_GEMM_DEEP", "0"), "M": M, "N": N, "K": K, "act": act,
                          "residual": res, "ms": round(t, 4), "tflops": round(2.0 * M * N * K / t / 1e9, 1),
                          "torch_addmm_tf32_ms": round(tl, 4)}))
    from openess_b200.models import maskclip_model as mm
    torch.manual_seed(1205)
    v = mm.maskClipFeatureExtractor(None, None, 11, None).cuda().eval()
    img = torch.rand(8, 3, 440, 640, device="cuda")
    print(json.dumps({"op": "maskclip_vit_b16_fwd", "B": 8, "deep": os.environ.get("OESS_GEMM_DEEP", "0"),
                      "ms": round(timeit(lambda: v(img), iters=10, warm=3), 3)}))


if __name__ == "__main__":
    main()
