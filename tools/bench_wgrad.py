"""Timing of the tcgen05 backward-weight kernel (oess_conv2d_wgrad_nhwc_tf32) at the SemSegE2VID decoder's shapes (B = 4)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("B", 4))
shapes = [(256, 256, 55, 80, 3), (256, 128, 55, 80, 3), (256, 128, 110, 160, 3), (128, 128, 110, 160, 3), (128, 64, 220, 320, 3),
          (64, 64, 220, 320, 3), (64, 32, 440, 640, 3)]
for (ci, co, H, W, k) in shapes:
    x = torch.randn(B, ci, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, co, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        ops.conv2d_wgrad(x, dy, k, padding=1)
    torch.cuda.synchronize()
    with _lib.profile() as prof:
        for _ in range(10):
            ops.conv2d_wgrad(x, dy, k, padding=1)
    ms = sum(v[1] for kk, v in prof.kernels.items() if "wgrad" in kk) / 10
    fl = 2.0 * B * H * W * ci * co * k * k
    print(json.dumps({"wgrad": f"{ci}->{co} {k}x{k} @{H}x{W}", "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                      "kernels": {kk: round(v[1] / 10, 4) for kk, v in prof.kernels.items()}}))
