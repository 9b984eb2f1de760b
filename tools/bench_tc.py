#!/usr/bin/env python
"""Times the tensor-core kernels against torch's library path (cuBLAS / cuDNN TF32) on the path's real shapes."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import ops  # noqa: E402


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
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    rows = []
    for (M, N, K) in [(8 * 17600, 256, 2048), (8 * 1121, 2304, 768), (8 * 1121, 768, 3072), (8 * 281600, 256, 32)]:
        a = torch.randn(M, K, device="cuda")
        b = torch.randn(N, K, device="cuda")
        out = torch.empty(M, N, device="cuda")
        t_ours = timeit(lambda: ops.gemm_tf32(a, b, None, out))
        t_lib = timeit(lambda: torch.matmul(a, b.t(), out=out))
        fl = 2.0 * M * N * K
        rows.append({"op": "gemm_tf32", "M": M, "N": N, "K": K, "ms": t_ours, "tflops": fl / t_ours / 1e9,
                     "torch_tf32_ms": t_lib, "torch_tflops": fl / t_lib / 1e9})
        print(json.dumps(rows[-1]))
        del a, b, out


def convlstm():
    """One ConvLSTM step at the three E2VID encoder levels of a DSEC batch (B = 8, 440 x 640 input)."""
    import torch.nn.functional as F
    from openess_b200 import losses
    for (C, H, W) in [(64, 220, 320), (128, 110, 160), (256, 55, 80)]:
        B = 8
        cl = torch.channels_last
        x = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=cl)
        h = torch.tanh(torch.randn(B, C, H, W, device="cuda")).contiguous(memory_format=cl)
        c = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=cl)
        wgt = (torch.randn(4 * C, 2 * C, 3, 3, device="cuda") / (18 * C) ** 0.5)
        bias = torch.randn(4 * C, device="cuda") * 0.1
        wp, bp = ops.convlstm_pack(wgt, bias, C)
        wcl = wgt.contiguous(memory_format=cl)

        def lib_path():
            gates = F.conv2d(torch.cat((x, h), 1), wcl, bias, padding=1)
            return losses.convlstm_gates(gates, c)

        def lib_path_nchw():
            gates = F.conv2d(torch.cat((x.contiguous(), h.contiguous()), 1), wgt, bias, padding=1)
            return losses.convlstm_gates(gates, c.contiguous())

        t_ours = timeit(lambda: ops.convlstm_step(x, (h, c), wp, bp))
        xb, wpb = x.bfloat16(), wp.bfloat16()
        hs = h.clone()
        hs._oess_bf16 = h.bfloat16()
        t_bf16 = timeit(lambda: ops.convlstm_step_bf16(xb, (hs, c), wpb, bp))
        t_lib = timeit(lib_path)
        t_lib2 = timeit(lib_path_nchw)
        fl = 2.0 * B * H * W * 4 * C * 18 * C
        print(json.dumps({"op": "convlstm_step", "B": B, "C": C, "H": H, "W": W, "ms": t_ours, "tflops": fl / t_ours / 1e9,
                          "ms_bf16_operands": t_bf16, "tflops_bf16_operands": fl / t_bf16 / 1e9,
                          "torch_cudnn_tf32_channels_last_plus_fused_gates_ms": t_lib, "torch_tflops": fl / t_lib / 1e9,
                          "torch_cudnn_tf32_nchw_plus_fused_gates_ms": t_lib2}))


def e2vid_step():
    """One recurrent step of the latent-only E2VID encoder on a DSEC batch (B = 8, 5 x 440 x 640), real width."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from seeded_weights import seeded_state_dict
    from openess_b200.e2vid.model import model as mm
    cfg = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
           'base_num_channels': 32, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}
    m = mm.E2VIDRecurrent(cfg, latent_only=True)
    m.load_state_dict(seeded_state_dict(m, 1205), strict=True)
    m = m.eval().cuda().fold_bn()
    x = torch.randn(8, 5, 440, 640, device="cuda")
    res = {}
    for use_tc in (True, False):
        mm.USE_TENSOR_CORES = use_tc
        with torch.no_grad():
            _, st, _ = m(x, None)

            def step():
                m(x, st)
            res[use_tc] = timeit(step, iters=10)
    mm.USE_TENSOR_CORES = True
    print(json.dumps({"op": "e2vid_latent_only_step", "B": 8, "H": 440, "W": 640, "ms_tensor_core_path": res[True],
                      "ms_cudnn_tf32_path": res[False], "frames_per_s_tensor_core_path": 8 / res[True] * 1e3}))


def teacher():
    """Forward of the dilated ResNet-50 teacher encoder on a DSEC batch (frames 3 x 440 x 640), train-mode BN (the
    trainers' operating point) and eval-mode (folded BN)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from seeded_weights import seeded_state_dict
    from openess_b200.models import image_model as im
    m = im.DilationFeatureExtractor()
    m.load_state_dict(seeded_state_dict(m, 77), strict=True)
    m = m.cuda()
    for B in (2, 8):
        x = torch.rand(B, 3, 440, 640, device="cuda")
        for mode in ("train", "eval"):
            m.train(mode == "train")
            res = {}
            for use_tc in (True, False):
                im.USE_TENSOR_CORES = use_tc
                with torch.no_grad():
                    res[use_tc] = timeit(lambda: m.encoder(x), iters=8, warm=4)   # warm >= 3: the third call captures the CUDA graph
            im.USE_TENSOR_CORES = True
            fl = 845.1e9 * B
            print(json.dumps({"op": "teacher_r50_dilated_encoder_fwd", "B": B, "bn": mode, "ms_tensor_core_path": res[True],
                              "tflops": fl / res[True] / 1e9, "ms_torch_cudnn_tf32_nchw": res[False],
                              "torch_tflops": fl / res[False] / 1e9}))
        del x
        torch.cuda.empty_cache()


def semseg():
    """SemSegE2VID forward (validation / linear-probing path) on a DSEC batch of E2VID latents, real width, K = 11."""
    from openess_b200.models import style_networks as sn
    m = sn.SemSegE2VID(input_c=256, output_c=11, skip_connect=True, skip_type='concat', text_embeddings_path=None).cuda().eval()
    B, H, W = 8, 440, 640
    lat = {8: torch.randn(B, 256, H // 8, W // 8, device="cuda"), 4: torch.randn(B, 128, H // 4, W // 4, device="cuda"),
           2: torch.randn(B, 64, H // 2, W // 2, device="cuda"), 1: torch.randn(B, 32, H, W, device="cuda")}
    res = {}
    for use_tc in (True, False):
        sn.USE_TENSOR_CORES = use_tc
        with torch.no_grad():
            res[use_tc] = timeit(lambda: m.forward_pooled(lat, torch.zeros(B, H, W, dtype=torch.int64, device="cuda"), 100, 800)[0][1],
                                 iters=5, warm=2)
    sn.USE_TENSOR_CORES = True
    fl = 175.0e9 * B
    print(json.dumps({"op": "semseg_e2vid_fwd_logits", "B": B, "ms_tensor_core_path": res[True], "tflops": fl / res[True] / 1e9,
                      "ms_torch_cudnn_tf32": res[False], "torch_tflops": fl / res[False] / 1e9}))


def deeplab():
    """BASELINE config 4: DeepLabv3-R50 head fine-tune step (frozen backbone, K = 11) forward + backward, 440 x 640."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from seeded_weights import seeded_state_dict
    from openess_b200.models import deeplabv3 as dl
    m = dl.deeplabv3_resnet50(num_classes=11, text_embeddings_path=None, output_stride=32, pretrained_backbone='',
                              if_finetuning=True, frozen_backbone=True)
    m.load_state_dict(seeded_state_dict(m, 4), strict=True)
    m = m.cuda().train()
    B = 4
    x = torch.rand(B, 3, 440, 640, device="cuda")

    def step():
        m.zero_grad(set_to_none=True)
        lo, fe = m(x)
        (lo.square().mean() + fe.square().mean()).backward()

    res = {}
    for mode in ("own", "torch"):
        dl.USE_TENSOR_CORES = mode == "own"
        res[mode] = timeit(step, iters=5, warm=2)
    dl.USE_TENSOR_CORES = True
    print(json.dumps({"op": "deeplabv3_r50_head_finetune_fwd_bwd", "B": B, "ms_own_kernels": res["own"],
                      "ms_torch_cudnn_tf32": res["torch"], "samples_per_s_own": B / res["own"] * 1e3}))


def config2():
    """BASELINE config 2 modules named by the north star: ResNet-18 (models/_resnet.py) and the MaskCLIP ViT-B/16
    (models/maskclip_model.py) on a DSEC batch of frames (3 x 440 x 640), own kernels against torch's TF32 library path of
    the same arithmetic for the ResNet (its torch formulation on cuDNN); the ViT mirror has no torch formulation in the
    package (and tools/ may not import oracle/), so it is reported with its per-kernel breakdown only."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    from seeded_weights import seeded_state_dict
    from openess_b200 import _lib
    from openess_b200.models._resnet import resnet18
    from openess_b200.models import maskclip_model as mm
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    m = resnet18(pretrained='')
    m.load_state_dict(seeded_state_dict(m, 18), strict=True)
    m = m.cuda()
    for p in m.parameters():
        p.requires_grad = False
    for B in (8, 32):
        x = torch.rand(B, 3, 440, 640, device="cuda")
        xcl = x.contiguous(memory_format=torch.channels_last)
        for mode in ("train", "eval"):
            m.train(mode == "train")
            with torch.no_grad():
                t_own = timeit(lambda: m(x), iters=10, warm=3)
                t_lib = timeit(lambda: m.forward_torch(x), iters=10, warm=3)
                m_cl = m.to(memory_format=torch.channels_last)
                t_lib_cl = timeit(lambda: m_cl.forward_torch(xcl), iters=10, warm=3)
                m.to(memory_format=torch.contiguous_format)
            fl = 20.5e9 * B
            print(json.dumps({"op": "resnet18_fwd", "B": B, "bn": mode, "ms_own_kernels": t_own, "tflops": fl / t_own / 1e9,
                              "ms_torch_cudnn_tf32_nchw": t_lib, "ms_torch_cudnn_tf32_nhwc": t_lib_cl}))
        del x, xcl
    torch.manual_seed(1205)
    v = mm.maskClipFeatureExtractor(None, None, 11, None)
    with torch.no_grad():
        v.encoder.pos_embed.normal_(0, 0.02)
        v.encoder.cls_token.normal_(0, 0.02)
    v = v.cuda().eval()
    for B in (1, 8):
        img = torch.rand(B, 3, 440, 640, device="cuda")
        t_own = timeit(lambda: v(img), iters=10, warm=3)
        with _lib.profile() as prof:
            v(img)
        T = 1121
        fl = B * (12 * (2 * T * 768 * 9216 + 4 * T * T * 768) + 2 * (T - 1) * 768 * 768 + 2 * T * 768 * (3 * 768 + 6144))
        print(json.dumps({"op": "maskclip_vit_b16_fwd", "B": B, "ms_own_kernels": t_own, "tflops": fl / t_own / 1e9,
                          "kernel_ms": {k: round(val[1], 4) for k, val in prof.kernels.items()}}))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


if __name__ == "__main__":
    if "--config2" in sys.argv:
        config2()
        sys.exit(0)
    if "--teacher" in sys.argv:
        teacher()
        semseg()
        deeplab()
        sys.exit(0)
    main()
    convlstm()
    e2vid_step()
