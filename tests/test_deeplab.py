"""DeepLabv3-ResNet-50 mirror (row a12) against goldens produced by the REFERENCE deeplabv3_resnet50 built the way
BASELINE config 4 builds it (K = 11, if_finetuning + frozen_backbone; oracle/make_golden_models.py --deeplab; weights from
tests/seeded_weights.py).  Tolerances: torch formulation on the CPU 3e-4 relative to the output scale; tensor-core paths
within 2x of torch's own cuDNN-TF32 deviation from the fp32 golden (train mode: a batch-statistics network amplifies
TF32 rounding; the comparison class is what the reference's GPU run does), eval mode 2e-2 of the output scale."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from seeded_weights import seeded_state_dict


def _build():
    from openess_b200.models.deeplabv3 import deeplabv3_resnet50
    z = load_golden("deeplab_r50")
    m = deeplabv3_resnet50(num_classes=11, text_embeddings_path=None, output_stride=32, pretrained_backbone='',
                           if_finetuning=True, frozen_backbone=True)
    sd = seeded_state_dict(m, int(z["seed"]))
    assert len(sd) == int(z["nkeys"]) and sum(p.numel() for p in m.parameters()) == int(z["nparams"])
    m.load_state_dict(sd, strict=True)
    m.classifier.ASPP.project[3].p = 0.0
    return z, m


def _sub(lo, fe):
    return lo.detach()[:, :, ::2, ::2].cpu().numpy(), fe.detach()[:, ::8, ::4, ::4].cpu().numpy()


def test_deeplab_mirror_structure_and_cpu_forward_backward():
    z, m = _build()
    assert not any(p.requires_grad for p in m.backbone.parameters()) and all(p.requires_grad for p in m.classifier.parameters())
    x = torch.from_numpy(z["x"])
    m.eval()
    with torch.no_grad():
        lo, fe = _sub(*m(x))
    np.testing.assert_allclose(lo, z["eval_logits_sub"], atol=3e-4 * float(np.abs(z["eval_logits_sub"]).max()))
    np.testing.assert_allclose(fe, z["eval_feats_sub"], atol=3e-4 * float(np.abs(z["eval_feats_sub"]).max()))
    m.train()
    lt, ft = m(x)
    (lt.square().mean() + ft.square().mean()).backward()
    lo, fe = _sub(lt, ft)
    np.testing.assert_allclose(lo, z["train_logits_sub"], atol=3e-4 * float(np.abs(z["train_logits_sub"]).max()))
    np.testing.assert_allclose(m.classifier.text_embeddings.grad.numpy(), z["grad_text"],
                               atol=1e-3 * float(np.abs(z["grad_text"]).max()))
    assert not any(p.grad is not None for p in m.backbone.parameters()) and not bool(z["backbone_has_grad"])
    assert m.classifier.pixel_feature.weight.grad is None          # unused parameter (deeplabv3.py:94)


@pytest.mark.gpu
def test_deeplab_frozen_backbone_tensor_cores_train_and_eval():
    from openess_b200 import _lib
    z, m = _build()
    m = m.cuda()
    x = torch.from_numpy(z["x"]).cuda()
    # ---- eval + no-grad (val_step / test.py): backbone AND head on the tensor cores, BN folded
    m.eval()
    with _lib.profile() as prof:
        with torch.no_grad():
            lo, fe = _sub(*m(x))
    assert prof.kernels["tc_conv2d"][0] == 53 + 4 + 1 + 1 and "bn_apply" not in prof.kernels
    for got, key in ((lo, "eval_logits_sub"), (fe, "eval_feats_sub")):
        assert float(np.abs(got - z[key]).max()) < 2e-2 * float(np.abs(z[key]).max()), key
    # ---- train mode (fine-tuning step, BASELINE config 4): frozen backbone on the tensor cores with batch-statistics BN,
    #      trainable head as conv_bn_autograd blocks (6 more tcgen05 convs forward; dgrad / wgrad / BN Jacobian backward)
    m.train()
    n0 = _lib.launch_count()
    with _lib.profile() as prof:
        lt, ft = m(x)
    assert prof.kernels["tc_conv2d"][0] == 53 + 6 and "bn_stats" not in prof.kernels and prof.kernels["bn_apply"][0] == 53 + 6
    n1 = _lib.launch_count()
    (lt.square().mean() + ft.square().mean()).backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n1 >= 6 * 3 + 2          # per block: BN bwd stats + apply + wgrad (+ dgrad past the first)
    np.testing.assert_allclose(m.classifier.text_embeddings.grad.cpu().numpy(), z["grad_text"],
                               atol=0.25 * float(np.abs(z["grad_text"]).max()))
    assert m.classifier.pixel_feature.weight.grad is None
    assert m.classifier.text_embeddings.grad is not None and not any(p.grad is not None for p in m.backbone.parameters())
    err_tc = np.abs(_sub(lt, ft)[0] - z["train_logits_sub"])
    # noise class: the same mirror through torch's default cuDNN-TF32 convolutions
    from openess_b200.models import deeplabv3 as dl
    m2 = _build()[1].cuda().train()
    dl.USE_TENSOR_CORES = False
    torch.backends.cudnn.allow_tf32 = True
    try:
        with torch.no_grad():
            l2, _ = m2(x)
    finally:
        torch.backends.cudnn.allow_tf32 = False
        dl.USE_TENSOR_CORES = True
    err_lib = np.abs(l2[:, :, ::2, ::2].cpu().numpy() - z["train_logits_sub"])
    scale = float(np.abs(z["train_logits_sub"]).max())
    print("deeplab train-mode logits (scale %.2f): tensor-core backbone mean |err| %.3e, torch cuDNN-TF32 %.3e"
          % (scale, err_tc.mean(), err_lib.mean()))
    assert err_tc.mean() < 2.0 * err_lib.mean() + 1e-4 * scale
    np.testing.assert_allclose(m.state_dict()["backbone.layer4.2.bn3.running_mean"].cpu().numpy(), z["rm_l4"], atol=5e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("C,Cout,k,p,d,res,relu", [(64, 64, 3, 1, 1, False, True), (256, 64, 1, 0, 1, False, True),
                                                   (32, 128, 3, 2, 2, True, True), (64, 32, 3, 1, 1, True, False)])
def test_conv_batchnorm_autograd_vs_torch(C, Cout, k, p, d, res, relu):
    """conv -> train-mode BatchNorm (-> + residual) (-> ReLU), forward and backward on hand-written kernels, against torch
    autograd in float64 (ReLU active set taken from the kernel's forward, see test_conv_instancenorm_autograd_vs_torch)."""
    import torch.nn.functional as F
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C + Cout + k)
    B, H, W = 2, 12, 20
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    conv_ref = torch.nn.Conv2d(C, Cout, k, padding=p, dilation=d, bias=False).double()
    bn_ref = torch.nn.BatchNorm2d(Cout).double().train()
    with torch.no_grad():
        conv_ref.weight.copy_(torch.randn(conv_ref.weight.shape, generator=g, dtype=torch.float64) / (C * k * k) ** 0.5)
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.normal_(0, 0.3)
    r = torch.randn(B, Cout, H, W, generator=g, dtype=torch.float64, requires_grad=True) if res else None
    conv = torch.nn.Conv2d(C, Cout, k, padding=p, dilation=d, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(Cout).cuda().train()
    conv.load_state_dict({k_: v.float() for k_, v in conv_ref.state_dict().items()})
    bn.load_state_dict({k_: (v.float() if v.is_floating_point() else v) for k_, v in bn_ref.state_dict().items()})
    xg = x.detach().float().cuda().requires_grad_(True)
    rg = r.detach().float().cuda().requires_grad_(True) if res else None
    yg = ops.conv_bn_autograd(xg, conv, bn, residual=rg, relu=relu)
    y = bn_ref(conv_ref(x))
    if res:
        y = y + r
    if relu:
        y = y * (yg.detach().cpu() > 0).double()
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    yg.backward(dy.float().cuda())
    assert float((yg.detach().cpu().double() - y.detach()).abs().max()) < 1e-2
    for got, ref, name in ((xg.grad, x.grad, "dx"), (conv.weight.grad, conv_ref.weight.grad, "dW"),
                           (bn.weight.grad, bn_ref.weight.grad, "dgamma"), (bn.bias.grad, bn_ref.bias.grad, "dbeta")):
        err = float((got.cpu().double() - ref).abs().max())
        assert err < 1e-2 * float(ref.abs().max()) + 1e-4, (name, err, float(ref.abs().max()))
    if res:
        assert float((rg.grad.cpu().double() - r.grad).abs().max()) < 1e-6
    torch.testing.assert_close(bn.running_var.cpu().double(), bn_ref.running_var, atol=1e-3, rtol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,h,w,H,W", [(2, 11, 14, 20, 440, 640), (1, 5, 7, 9, 30, 23), (1, 3, 16, 16, 16, 16), (2, 4, 9, 5, 4, 3)])
def test_bilinear_resize_matches_torch_forward_and_backward(B, C, h, w, H, W):
    """oess_bilinear_resize_planes (+ separable gather backward) against F.interpolate(bilinear, align_corners=False) and its
    autograd backward (deeplabv3.py:53-56): up- and down-scaling, non-integer ratios, identity."""
    import torch.nn.functional as F
    from openess_b200 import ops
    g = torch.Generator().manual_seed(h * 31 + W)
    x = torch.randn(B, C, h, w, generator=g).cuda().requires_grad_(True)
    wgt = torch.randn(B, C, H, W, generator=g).cuda()
    out = ops.bilinear_resize(x, (H, W))
    ref = F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    gx, = torch.autograd.grad((out * wgt).sum(), x)
    rx, = torch.autograd.grad((ref * wgt).sum(), x)
    torch.testing.assert_close(gx, rx, rtol=1e-4, atol=1e-4 * float(rx.abs().max()))
