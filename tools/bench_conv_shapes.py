"""Per-layer timing of the teacher's conv + train-mode BatchNorm blocks (distinct shapes of the dilated ResNet-50 at a DSEC batch):
conv kernel (tcgen05, fused statistics) and BatchNorm apply separately, with the FLOP rate and the HBM bytes each moves."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("B", 4))
shapes = [  # (Cin, Cout, H, W, k, stride, pad, dil)
    (64, 64, 110, 160, 1, 1, 0, 1), (64, 64, 110, 160, 3, 1, 1, 1), (64, 256, 110, 160, 1, 1, 0, 1), (256, 64, 110, 160, 1, 1, 0, 1),
    (256, 128, 110, 160, 1, 1, 0, 1), (128, 128, 110, 160, 3, 2, 1, 1), (128, 512, 55, 80, 1, 1, 0, 1), (512, 128, 55, 80, 1, 1, 0, 1),
    (128, 128, 55, 80, 3, 1, 1, 1), (512, 256, 55, 80, 1, 1, 0, 1), (256, 256, 55, 80, 3, 1, 2, 2), (256, 1024, 55, 80, 1, 1, 0, 1),
    (1024, 256, 55, 80, 1, 1, 0, 1), (1024, 512, 55, 80, 1, 1, 0, 1), (512, 512, 55, 80, 3, 1, 4, 4), (512, 2048, 55, 80, 1, 1, 0, 1),
    (2048, 512, 55, 80, 1, 1, 0, 1),
]
if os.environ.get("SHAPE"):                                   # e.g. SHAPE=11 : one shape only (for an ncu capture)
    shapes = [shapes[int(os.environ["SHAPE"])]]
dev = torch.device("cuda")
for (ci, co, H, W, k, s, p, d) in shapes:
    x = torch.randn(B, ci, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, k, k, device=dev) * 0.05
    wp = ops.conv2d_pack(w)
    bn = torch.nn.BatchNorm2d(co).to(dev).train()
    for _ in range(3):
        y = ops.conv_bn_train(x, wp, None, k, s, p, d, bn, relu=True)
    torch.cuda.synchronize()
    with _lib.profile() as prof:
        for _ in range(10):
            y = ops.conv_bn_train(x, wp, None, k, s, p, d, bn, relu=True)
    Ho, Wo = y.shape[2], y.shape[3]
    ms_conv = prof.kernels["tc_conv2d"][1] / 10
    ms_bn = (prof.kernels["bn_apply"][1] + prof.kernels.get("bn_finalize", (0, 0.0))[1]) / 10
    fl = 2.0 * B * Ho * Wo * co * ci * k * k
    by = 4.0 * B * (H * W * ci + Ho * Wo * co)
    print(json.dumps({"conv": f"{ci}->{co} {k}x{k} s{s} d{d} @{H}x{W}", "ms_conv": round(ms_conv, 4), "tflops": round(fl / ms_conv / 1e9, 1),
                      "conv_GBs": round(by / ms_conv / 1e6, 0), "ms_bn_apply": round(ms_bn, 4),
                      "bn_GBs": round(8.0 * B * Ho * Wo * co / ms_bn / 1e6, 0)}))
