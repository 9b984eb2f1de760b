"""Dilated ResNet-50 teacher mirror (row a13) against goldens produced by the REFERENCE DilationFeatureExtractor in
train mode (oracle/make_golden_models.py --teacher; weights regenerated from tests/seeded_weights.py).

Tolerances: torch formulation (CPU, fp32, same ops as the reference) 2e-4; tensor-core formulation: TF32 operands through
52 convolutions, each re-normalised by batch-statistics BatchNorm -> 3e-2 abs on features whose max |value| is ~10
(measured value printed), 2e-2 on the unit-norm output features."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from seeded_weights import seeded_state_dict


def _build():
    from openess_b200.models.image_model import DilationFeatureExtractor
    z = load_golden("teacher_r50")
    m = DilationFeatureExtractor()
    m.load_state_dict(seeded_state_dict(m, int(z["seed"])), strict=True)
    return z, m.train()


def test_teacher_mirror_structure_and_cpu_forward():
    z, m = _build()
    assert len(m.state_dict()) == 320 and sum(p.numel() for p in m.parameters()) == 24032576
    assert not any(p.requires_grad for p in m.encoder.parameters()) and all(p.requires_grad for p in m.decoder.parameters())
    y = m(torch.from_numpy(z["x"]))
    np.testing.assert_allclose(y.detach()[:, ::8, ::4, ::4].numpy(), z["y_sub"], atol=2e-4)
    sd = m.state_dict()
    np.testing.assert_allclose(sd["encoder.layer4.2.bn3.running_mean"].numpy(), z["rm_l4"], atol=1e-4)
    assert int(sd["encoder.layer3.5.bn2.num_batches_tracked"]) == int(z["nbt"]) == 1
    y.square().mean().backward()                             # decoder trains, encoder frozen
    assert m.decoder[0].weight.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_teacher_tensor_core_train_mode_vs_reference_golden(dtype, monkeypatch):
    """bf16 (opt-in, OESS_TEACHER_DTYPE=bf16): bfloat16 conv operands after the stem, fp32 accumulation / statistics /
    residuals.  Stated tolerance against the fp32 golden: 8x the error of torch's cuDNN-TF32 run of the same network on the
    layer4 features (measured 6x: 2 fewer mantissa bits through 52 batch-renormalised convs of a random network),
    3e-2 mean / 5e-1 max on the unit-norm output features."""
    from openess_b200 import _lib
    from openess_b200.models import image_model as im
    monkeypatch.setattr(im, "TEACHER_BF16", dtype == "bf16")
    z, m = _build()
    m = m.cuda()
    x = torch.from_numpy(z["x"]).cuda()
    with _lib.profile() as prof:
        feats = m.encoder(x)
    if dtype == "bf16":
        assert prof.kernels["tc_conv2d"][0] == 1 and prof.kernels["tc_conv2d_bf16"][0] == 52
    else:
        assert prof.kernels["tc_conv2d"][0] == 53
    assert "bn_stats" not in prof.kernels and prof.kernels["bn_apply"][0] == 53
    k = 8.0 if dtype == "bf16" else 2.0
    assert prof.kernels["maxpool3x3s2_nhwc"][0] == 1        # stem conv + pool on own kernels too
    err_f = np.abs(feats[:, ::16].cpu().numpy() - z["feats_sub"])
    # the noise class this has to stay in: torch's own default GPU arithmetic for the same network (cuDNN TF32
    # convolutions, what the reference's GPU run does) against the same CPU fp32 golden
    m2 = _build()[1].cuda()
    torch.backends.cudnn.allow_tf32 = True
    try:
        with torch.no_grad():
            feats_lib = m2.encoder.forward_torch(x)
    finally:
        torch.backends.cudnn.allow_tf32 = False
    err_lib = np.abs(feats_lib[:, ::16].cpu().numpy() - z["feats_sub"])
    print("teacher features (max |value| %.1f): %s tensor-core path max / mean |err| %.3e / %.3e; torch cuDNN-TF32 path %.3e / %.3e"
          % (float(z["feats_absmax"]), dtype, err_f.max(), err_f.mean(), err_lib.max(), err_lib.mean()))
    assert err_f.mean() < k * err_lib.mean() + 1e-4 and err_f.max() < k * err_lib.max() + 1e-3
    sd = m.state_dict()                                      # running statistics updated exactly once, like the reference
    np.testing.assert_allclose(sd["encoder.layer4.2.bn3.running_mean"].cpu().numpy(), z["rm_l4"], atol=2e-3)
    np.testing.assert_allclose(sd["encoder.layer1.0.bn1.running_var"].cpu().numpy(), z["rv_l1"], rtol=2e-3)
    np.testing.assert_allclose(sd["encoder.layer2.0.downsample.1.running_var"].cpu().numpy(), z["rv_ds"], rtol=2e-3)
    assert int(sd["encoder.layer3.5.bn2.num_batches_tracked"]) == 1
    m.load_state_dict({k: v.cuda() for k, v in seeded_state_dict(m, int(z["seed"])).items()}, strict=True)
    y = m(x)                                                 # full forward: encoder (tensor cores) + trainable decoder
    err_y = np.abs(y.detach()[:, ::8, ::4, ::4].cpu().numpy() - z["y_sub"])
    print("teacher unit-norm output features (%s): max / mean |err| %.3e / %.3e" % (dtype, err_y.max(), err_y.mean()))
    assert (err_y.mean() < 3e-2 and err_y.max() < 5e-1) if dtype == "bf16" else (err_y.mean() < 3e-3 and err_y.max() < 5e-2)
    y.square().mean().backward()
    assert m.decoder[0].weight.grad is not None and float(m.decoder[0].weight.grad.abs().max()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_teacher_eval_mode_folded_bn_vs_torch(dtype, monkeypatch):
    from openess_b200 import _lib
    from openess_b200.models import image_model as im
    monkeypatch.setattr(im, "TEACHER_BF16", dtype == "bf16")
    z, m = _build()
    m = m.cuda().eval()
    x = torch.from_numpy(z["x"]).cuda()
    with torch.no_grad():
        ref = m.encoder.forward_torch(x)                     # cuDNN fp32 (TF32 disabled in conftest)
        with _lib.profile() as prof:
            out = m.encoder(x)
    nconv = prof.kernels["tc_conv2d"][0] + (prof.kernels["tc_conv2d_bf16"][0] if dtype == "bf16" else 0)
    assert nconv == 53 and "bn_apply" not in prof.kernels    # BN folded: no BN pass at all
    scale = float(ref.abs().max())
    print("teacher eval-mode features (%s): max |err| / max |value| %.3e" % (dtype, float((out - ref).abs().max()) / scale))
    assert float((out - ref).abs().max()) < (6e-2 if dtype == "bf16" else 2e-2) * scale


@pytest.mark.gpu
@pytest.mark.parametrize("C,training,res,relu", [(64, True, False, True), (256, True, True, True), (2048, True, False, False),
                                                 (512, False, True, True), (16, True, False, False)])
def test_batchnorm_nhwc_vs_torch(C, training, res, relu):
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C)
    B, H, W = 2, 9, 13
    x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.5).cuda()
    r = torch.randn(B, C, H, W, generator=g).cuda() if res else None
    bn_ref = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.normal_(0, 0.3)
        bn_ref.running_mean.normal_(0, 0.2)
        bn_ref.running_var.uniform_(0.5, 2.0)
    bn = torch.nn.BatchNorm2d(C).cuda()
    bn.load_state_dict(bn_ref.state_dict())
    bn_ref.train(training)
    bn.train(training)
    with torch.no_grad():
        ref = bn_ref(x)
        if res:
            ref = ref + r
        if relu:
            ref = ref.relu()
    y = ops.batchnorm_nhwc_(x.clone().contiguous(memory_format=torch.channels_last), bn, residual=r, relu=relu)
    torch.testing.assert_close(y, ref, atol=2e-5, rtol=2e-5)
    torch.testing.assert_close(bn.running_mean, bn_ref.running_mean, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(bn.running_var, bn_ref.running_var, atol=1e-6, rtol=1e-5)
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)


@pytest.mark.gpu
@pytest.mark.parametrize("B,h,w,S", [(2, 11, 16, 12), (1, 5, 9, 7)])
def test_upnorm_pool_fused_vs_torch_formulation(B, h, w, S):
    """oess_upnorm_pool (fwd + bwd) against Upsample(x4, bilinear, align_corners=True) + F.normalize + the one-hot
    sparse pooling of pretrain_trainer.py:446-463, in float64."""
    import torch.nn.functional as F
    from openess_b200 import ops
    g = torch.Generator().manual_seed(h * 31 + w)
    C, H, W = 256, 4 * h, 4 * w
    d = torch.randn(B, C, h, w, generator=g).cuda().requires_grad_(True)
    sp = torch.randint(0, S, (B, H, W), generator=g).cuda()
    sp[:, :, : W // 2] = sp[:, :1, :1]                       # long runs of one id (the register-accumulation path)
    M = B * S
    q = ops.upnorm_pool(d, sp, S, M)
    wgt = torch.randn(M, C, generator=g).cuda()
    (q * wgt).sum().backward()
    dd = d.detach().double().requires_grad_(True)
    feat = F.normalize(F.interpolate(dd, scale_factor=4, mode="bilinear", align_corners=True), p=2, dim=1)
    ids = (torch.arange(0, B * S, S, device=sp.device)[:, None, None] + sp).flatten()
    onehot = torch.zeros(M, ids.numel(), dtype=torch.float64, device=sp.device)
    onehot[ids, torch.arange(ids.numel(), device=sp.device)] = 1
    qr = (onehot @ feat.permute(0, 2, 3, 1).flatten(0, 2)) / (onehot.sum(1, keepdim=True) + 1e-6)
    (qr * wgt.double()).sum().backward()
    torch.testing.assert_close(q.double(), qr, atol=2e-5, rtol=2e-5)
    torch.testing.assert_close(d.grad.double(), dd.grad, atol=2e-4 * float(dd.grad.abs().max()), rtol=1e-3)


@pytest.mark.gpu
def test_teacher_encoder_cuda_graph_equals_eager(monkeypatch):
    """The frozen tensor-core encoder forward as a CUDA graph (third call on): same features as the eager path for changing inputs,
    running statistics advance alike, and the graph is dropped when the weights are reloaded."""
    from openess_b200.models import image_model as im
    z, m = _build()
    m = m.cuda()
    _, ref = _build()
    ref = ref.cuda()
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        x = torch.rand(1, 3, 64, 96, generator=g).cuda()
        monkeypatch.setattr(im, "TEACHER_GRAPH", True)
        got = m.encoder(x).clone()
        monkeypatch.setattr(im, "TEACHER_GRAPH", False)
        want = ref.encoder(x)
        assert torch.equal(got, want), it
        assert (m.encoder._graphed().graph is not None) == (it >= 2)
    torch.testing.assert_close(m.encoder.layer4[2].bn3.running_mean, ref.encoder.layer4[2].bn3.running_mean)
    assert int(m.encoder.bn1.num_batches_tracked) == int(ref.encoder.bn1.num_batches_tracked) == 5
    m.load_state_dict({k: v.cuda() for k, v in seeded_state_dict(m, int(z["seed"]) + 1).items()}, strict=True)
    monkeypatch.setattr(im, "TEACHER_GRAPH", True)
    x = torch.rand(1, 3, 64, 96, generator=g).cuda()
    y1 = m.encoder(x).clone()
    assert m.encoder._graphed().graph is None                # new weights: eager again until re-captured
    monkeypatch.setattr(im, "TEACHER_GRAPH", False)
    ref.load_state_dict({k: v.cuda() for k, v in seeded_state_dict(ref, int(z["seed"]) + 1).items()}, strict=True)
    assert torch.equal(y1, ref.encoder(x))
