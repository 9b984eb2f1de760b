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
    assert prof.kernels["tc_conv2d"][0] == 52 + 4 + 1 + 1 and "bn_apply" not in prof.kernels
    for got, key in ((lo, "eval_logits_sub"), (fe, "eval_feats_sub")):
        assert float(np.abs(got - z[key]).max()) < 2e-2 * float(np.abs(z[key]).max()), key
    # ---- train mode (fine-tuning step): frozen backbone on the tensor cores with batch-statistics BN, head under autograd
    m.train()
    with _lib.profile() as prof:
        lt, ft = m(x)
    assert prof.kernels["tc_conv2d"][0] == 52 and "bn_stats" not in prof.kernels and prof.kernels["bn_apply"][0] == 52
    (lt.square().mean() + ft.square().mean()).backward()
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
