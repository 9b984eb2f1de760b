"""ResNet-18 mirror (SURVEY 8a row a14', models/_resnet.py:225-234) against goldens produced by the REFERENCE's own
`resnet18(pretrained='')` on CPU (oracle/make_golden_models.py --resnet18; weights regenerated from
tests/seeded_weights.py).  Tolerances: torch formulation (CPU fp32, same ops) 2e-4; tensor-core formulation (TF32
operands through 20 convolutions + the fc): compared with torch's own cuDNN-TF32 path on the same golden (printed)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from seeded_weights import seeded_state_dict


def _build():
    from openess_b200.models._resnet import resnet18
    z = load_golden("resnet18")
    m = resnet18(pretrained='')
    m.load_state_dict(seeded_state_dict(m, int(z["seed"])), strict=True)
    return z, m


def test_resnet18_structure_and_cpu_forward():
    z, m = _build()
    assert len(m.state_dict()) == int(z["nkeys"]) == 122 and sum(p.numel() for p in m.parameters()) == int(z["nparams"]) == 11689512
    m.eval()
    x = torch.from_numpy(z["x"])
    with torch.no_grad():
        np.testing.assert_allclose(m(x).numpy(), z["eval_logits"], atol=2e-4)
        np.testing.assert_allclose(m.forward_features(x).numpy(), z["eval_feats"], atol=2e-4)
    m.train()
    with torch.no_grad():
        np.testing.assert_allclose(m(x).numpy(), z["train_logits"], atol=2e-4)
    np.testing.assert_allclose(m.state_dict()["layer4.1.bn2.running_mean"].numpy(), z["rm_l4"], atol=1e-5)
    from openess_b200.models import _resnet
    with pytest.raises(RuntimeError):
        _resnet.resnet18(pretrained='imagenet')              # no network: loud, not silent


@pytest.mark.gpu
def test_pool_kernels_vs_torch():
    from openess_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    for shape in ((2, 64, 48, 80), (1, 8, 7, 9), (3, 12, 5, 4)):
        x = torch.randn(shape, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        assert torch.equal(ops.maxpool3x3s2_nhwc(x), torch.nn.functional.max_pool2d(x, 3, 2, 1))
        np.testing.assert_allclose(ops.global_avgpool_nhwc(x).cpu().numpy(), x.double().mean((2, 3)).cpu().numpy(), atol=1e-6)


@pytest.mark.gpu
def test_resnet18_tensor_core_vs_reference_golden():
    from openess_b200 import _lib
    z, m = _build()
    m = m.cuda().eval()
    x = torch.from_numpy(z["x"]).cuda()
    with torch.no_grad(), _lib.profile() as prof:
        logits = m(x)
    # 20 convs (stem, 16 block convs, 3 downsample) on the conv kernel, BN folded; own pools; fc on the GEMM
    assert prof.kernels["tc_conv2d"][0] == 20 and prof.kernels["maxpool3x3s2_nhwc"][0] == 1
    assert prof.kernels["global_avgpool_nhwc"][0] == 1 and prof.kernels["tc_gemm_tf32"][0] == 1
    with torch.no_grad():
        feats = m.forward_features(x)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            lib_logits = m.forward_torch(x)
        finally:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    e_own = np.abs(logits.cpu().numpy() - z["eval_logits"])
    e_lib = np.abs(lib_logits.cpu().numpy() - z["eval_logits"])
    e_f = np.abs(feats.cpu().numpy() - z["eval_feats"])
    print("resnet18 eval logits (max |value| %.2f): own max / mean |err| %.3e / %.3e; torch cuDNN-TF32 %.3e / %.3e; layer4 map %.3e"
          % (np.abs(z["eval_logits"]).max(), e_own.max(), e_own.mean(), e_lib.max(), e_lib.mean(), e_f.max()))
    # measured on B200: own 2.6e-3 / 6.7e-4 (max / mean), torch cuDNN-TF32 1.4e-3 / 3.2e-4 -- same class, margin for cuDNN's algorithm choice
    assert e_own.mean() < 3.0 * e_lib.mean() + 2e-4 and e_own.max() < 3.0 * e_lib.max() + 2e-3
    assert e_f.max() < 2e-2 * max(1.0, float(np.abs(z["eval_feats"]).max()))
    # train mode (how the OpenESS trainers run every network): batch statistics + running-stat update, once
    m.train()
    for p in m.parameters():
        p.requires_grad = False
    lt = m(x)
    e_t = np.abs(lt.cpu().numpy() - z["train_logits"])
    print("resnet18 train-mode logits: own max / mean |err| %.3e / %.3e" % (e_t.max(), e_t.mean()))
    assert e_t.max() < 3e-2 * max(1.0, float(np.abs(z["train_logits"]).max()))
    sd = m.state_dict()
    np.testing.assert_allclose(sd["layer4.1.bn2.running_mean"].cpu().numpy(), z["rm_l4"], atol=2e-3)
    np.testing.assert_allclose(sd["bn1.running_var"].cpu().numpy(), z["rv_stem"], rtol=2e-3, atol=1e-4)
    assert int(sd["layer2.0.downsample.1.num_batches_tracked"]) == int(z["nbt"]) == 1
