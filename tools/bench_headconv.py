"""Timing of the E2VID head convolution (5 -> 32 channels, 5 x 5, row-unfolded thin-input kernel) at a DSEC batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import ops  # noqa: E402

B, H, W = int(os.environ.get("B", 4)), 440, 640
x = torch.randn(B, 5, H, W, device="cuda")
w = torch.zeros(32, 8, 5, 5, device="cuda")
w[:, :5] = torch.randn(32, 5, 5, 5, device="cuda") * 0.1
wp = ops.conv2d_pack_rowunfold(w)
b = torch.zeros(32, device="cuda")
x8 = ops.planes_to_nhwc_padded_w(x, 8, 2)
for _ in range(3):
    y = ops.conv2d_rowunfold(x8, wp, b, 5, 5, W, relu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    y = ops.conv2d_rowunfold(x8, wp, b, 5, 5, W, relu=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"head conv B={B}: {ms:.4f} ms, output {y.numel() * 4 / ms / 1e6:.0f} GB/s")
