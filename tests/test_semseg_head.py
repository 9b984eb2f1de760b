"""SemSegE2VID mirror with the collapsed (fused) head (a11) against a golden produced by the reference class
(oracle/make_golden_models.py --semseg), and the pixel_linear kernels against torch convolutions.
The collapsed head is mathematically exact but rounds differently (32-term instead of 256/512-term dot products):
tolerance 1e-4 relative to the logits' scale."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def _build(z, dev="cpu"):
    from openess_b200.models.style_networks import SemSegE2VID
    m = SemSegE2VID(input_c=32, output_c=int(z["K"]), skip_connect=True, skip_type='concat', text_embeddings_path=None)
    m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}, strict=True)
    return m.to(dev)


def _latents(z, dev, grad=False):
    lat = {k: torch.from_numpy(z[f"lat{k}"]).to(dev) for k in (8, 4, 2, 1)}
    if grad:
        for k in (8, 4, 2):
            lat[k].requires_grad_(True)
    return lat


def test_semseg_mirror_state_dict_and_cpu_forward():
    z = load_golden("semseg_tiny")
    m = _build(z)
    out, x256 = m(_latents(z, "cpu"))
    scale = np.abs(z["logits"]).max()
    np.testing.assert_allclose(out[1].detach().numpy(), z["logits"], atol=1e-4 * scale)
    np.testing.assert_allclose(out[2].detach().numpy(), z["out2"], atol=1e-5)
    np.testing.assert_allclose(out[4].detach().numpy(), z["out4"], atol=1e-5)
    np.testing.assert_allclose(x256.detach().numpy(), z["x256"], atol=1e-5)
    assert sorted(out) == [1, 2, 4, 8]


@pytest.mark.gpu
@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 32, 11, 20, 28), (1, 4, 6, 16, 24), (3, 64, 64, 9, 13), (2, 11, 32, 7, 5)])
def test_pixel_linear_vs_torch(B, Cin, Cout, H, W):
    from openess_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Cin * 100 + Cout)
    x = torch.randn((B, Cin, H, W), generator=g).to(dev).requires_grad_(True)
    Wt = torch.randn((Cout, Cin), generator=g).to(dev).requires_grad_(True)
    b = torch.randn(Cout, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn((B, Cout, H, W), generator=g).to(dev)
    y = ops.pixel_linear(x, Wt, b)
    y.backward(gy)
    got = [t.grad.clone() for t in (x, Wt, b)]
    for t in (x, Wt, b):
        t.grad = None
    yr = torch.nn.functional.conv2d(x.double(), Wt.double()[:, :, None, None], b.double())
    yr.backward(gy.double())
    torch.testing.assert_close(y.double(), yr, atol=1e-5, rtol=1e-5)
    for a, t in zip(got, (x, Wt, b)):
        torch.testing.assert_close(a.double(), t.grad.double(), atol=2e-4, rtol=1e-5)
    y2 = ops.pixel_linear(x.detach(), Wt.detach())                 # no bias
    torch.testing.assert_close(y2.double(), torch.nn.functional.conv2d(x.detach().double(), Wt.detach().double()[:, :, None, None]),
                               atol=1e-5, rtol=1e-5)


@pytest.mark.gpu
def test_semseg_fused_head_forward_backward_golden():
    from openess_b200.models import style_networks as sn
    z = load_golden("semseg_tiny")
    dev = torch.device("cuda:0")
    m = _build(z, dev)
    lat = _latents(z, dev, grad=True)
    sp = torch.from_numpy(z["sp"]).to(dev)
    sn.TRAIN_ON_TENSOR_CORES = False          # strict fp32 formulation against the fp32 CPU golden (tolerances below)
    try:
        _golden_checks(z, m, lat, sp)
    finally:
        sn.TRAIN_ON_TENSOR_CORES = True


def _golden_checks(z, m, lat, sp):
    out, k = m.forward_pooled(lat, sp, int(z["S"]))
    scale = float(np.abs(z["logits"]).max())
    np.testing.assert_allclose(out[1].detach().cpu().numpy(), z["logits"], atol=2e-4 * scale)
    np.testing.assert_allclose(k.detach().cpu().numpy(), z["k"], atol=2e-4 * float(np.abs(z["k"]).max()))
    loss = out[1].square().mean() + 3.0 * k.square().mean()
    assert float(loss.detach()) == pytest.approx(float(z["loss"]), rel=2e-4)
    loss.backward()
    for kk in (8, 4, 2):
        ref = z[f"dlat{kk}"]
        np.testing.assert_allclose(lat[kk].grad.cpu().numpy(), ref, atol=3e-4 * float(np.abs(ref).max()))
    named = dict(m.named_parameters())
    checked = 0
    for key in z.files:
        if key.startswith("grad__"):
            ref = z[key]
            got = named[key[6:]].grad
            assert got is not None, key
            atol = 3e-4 * float(np.abs(ref).max()) + 1e-9
            wkey = key[:-4] + "weight"
            if key.endswith(".model.0.bias") and wkey in z.files:
                # a bias in front of an affine-free InstanceNorm has an exactly-zero gradient; the reference's
                # value is float round-off noise (5e-4 next to weight grads of 1e3), so only its scale is compared
                atol += 3e-6 * float(np.abs(z[wkey]).max())
            np.testing.assert_allclose(got.cpu().numpy(), ref, atol=atol)
            checked += 1
    assert checked >= 7
    # parameters that get no gradient in the reference forward get none here either (decoder_scale_5)
    for n in z["nograd"]:
        assert named[str(n)].grad is None
    # reference-signature forward on the GPU (materialises x_ch256 but not the 512-channel map)
    out2, x256 = m({k_: v.detach() for k_, v in lat.items()})
    np.testing.assert_allclose(x256.detach().cpu().numpy(), z["x256"], atol=2e-4)
    np.testing.assert_allclose(out2[1].detach().cpu().numpy(), out[1].detach().cpu().numpy(), atol=1e-5 * scale)


@pytest.mark.gpu
def test_semseg_trunk_tensor_cores_forward_only():
    """Validation / linear-probing path: conv + InstanceNorm (+ residual) (+ ReLU) of the whole trunk on the tcgen05 conv
    kernel with per-sample statistics in its epilogue.  Real width (input_c = 256, K = 11) against the same module's torch
    formulation in fp32 (cuDNN, TF32 disabled by conftest); tolerance: TF32 operands through 16 convs, each re-normalised
    by InstanceNorm -> 3e-2 of the output scale (measured value printed)."""
    from openess_b200 import _lib
    from openess_b200.models import style_networks as sn
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = sn.SemSegE2VID(input_c=256, output_c=11, skip_connect=True, skip_type='concat', text_embeddings_path=None).to(dev).eval()
    with torch.no_grad():
        m.text_embeddings.normal_(0, 0.3)
        for p in m.parameters():
            if p.ndim == 4:
                p.mul_(6.0)
    B, H, W = 2, 48, 80
    g = torch.Generator(device="cuda").manual_seed(5)
    lat = {8: torch.randn(B, 256, H // 8, W // 8, device=dev, generator=g), 4: torch.randn(B, 128, H // 4, W // 4, device=dev, generator=g),
           2: torch.randn(B, 64, H // 2, W // 2, device=dev, generator=g), 1: torch.randn(B, 32, H, W, device=dev, generator=g)}
    with torch.no_grad():
        sn.USE_TENSOR_CORES = False
        try:
            ref_out, ref_x = m(lat)
        finally:
            sn.USE_TENSOR_CORES = True
        with _lib.profile() as prof:
            out, x256 = m(lat)
    assert prof.kernels["tc_conv2d"][0] == 16 and prof.kernels["in_apply"][0] == 16      # 5 x 2 + 1 + 2 + 2 + 1 convs
    for k in (1, 2, 4):
        scale = float(ref_out[k].abs().max())
        err = float((out[k] - ref_out[k]).abs().max())
        print("SemSegE2VID tensor-core trunk: out[%d] max |err| %.3e of scale %.2f" % (k, err, scale))
        assert err < 3e-2 * scale
    assert float((x256 - ref_x).abs().max()) < 3e-2 * float(ref_x.abs().max())
    # under autograd with a trainable trunk every block is the differentiable tensor-core block: forward, backward-data and
    # backward-weight convolutions all launch tcgen05 kernels; gradients agree with the torch (cuDNN fp32) formulation
    sp = torch.randint(0, 20, (B, H, W), device=dev)
    grads = {}
    for mode in ("fp32", "cudnn_tf32", "tc"):            # torch fp32 / torch with cuDNN TF32 (its default) / own kernels
        sn.TRAIN_ON_TENSOR_CORES = mode == "tc"
        torch.backends.cudnn.allow_tf32 = mode == "cudnn_tf32"
        try:
            m.zero_grad(set_to_none=True)
            n0 = _lib.launch_count()
            with _lib.profile() as prof2:
                out3, k3 = m.forward_pooled(lat, sp, 20)
            n1 = _lib.launch_count()
            (out3[1].square().mean() + k3.square().mean()).backward()
            torch.cuda.synchronize()
            n2 = _lib.launch_count()
        finally:
            sn.TRAIN_ON_TENSOR_CORES = True
            torch.backends.cudnn.allow_tf32 = False
        assert (prof2.kernels.get("tc_conv2d", (0, 0))[0] == 16) == (mode == "tc")
        # backward kernels run on autograd's worker thread (the recorder is per host thread): count launches process-wide.
        # tensor-core path: per block in_bwd_stats + in_bwd_apply + wgrad (+ dgrad except for the first blocks' frozen inputs)
        if mode == "tc":
            assert n2 - n1 >= 16 * 3 + 14
        else:
            assert n2 - n1 < 16
        grads[mode] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    assert set(grads["tc"]) == set(grads["fp32"])
    # a 16-layer InstanceNorm / ReLU network amplifies operand rounding on the way back: the yardstick is what torch's own
    # default arithmetic (cuDNN TF32) does to the same gradients
    worst_tc = worst_lib = 0.0
    for n, gref in grads["fp32"].items():
        if n.endswith(".model.0.bias") or n.endswith(".model.3.bias"):
            continue                                  # bias in front of InstanceNorm: exactly-zero gradient, round-off only
        den = float(gref.abs().max()) + 1e-12
        worst_tc = max(worst_tc, float((grads["tc"][n] - gref).abs().max()) / den)
        worst_lib = max(worst_lib, float((grads["cudnn_tf32"][n] - gref).abs().max()) / den)
    print("SemSegE2VID training trunk: worst parameter-gradient deviation from fp32: own kernels %.3e, torch cuDNN-TF32 %.3e"
          % (worst_tc, worst_lib))
    assert worst_tc < 3.0 * worst_lib + 1e-3
    # linear probing: trunk frozen -> tensor cores even with grad enabled; only linear_probe trains (style_networks.py:169-170)
    mp = sn.SemSegE2VID(input_c=256, output_c=11, skip_connect=True, skip_type='concat', text_embeddings_path=None,
                        if_linear_probing=True).to(dev)
    with _lib.profile() as prof3:
        outp, _ = mp(lat)
        outp[1].mean().backward()
    assert prof3.kernels["tc_conv2d"][0] == 16 and mp.linear_probe.weight.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize("C,Cout,res,relu", [(64, 64, True, False), (32, 128, False, True), (16, 8, False, True)])
def test_conv_instancenorm_vs_torch(C, Cout, res, relu):
    import torch.nn.functional as F
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C * 3 + Cout)
    B, H, W = 3, 13, 21
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / (9 * C) ** 0.5
    b = torch.randn(Cout, generator=g)
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    ref = F.instance_norm(F.conv2d(x.double(), w.double(), b.double(), padding=1), eps=1e-5)
    if res:
        ref = ref + r.double()
    if relu:
        ref = ref.relu()
    y = ops.conv_in(x.cuda(), ops.conv2d_pack(w.cuda()), b.cuda(), 3, 1, 1, 1, residual=None if r is None else r.cuda(), relu=relu)
    assert float((y.cpu().double() - ref).abs().max()) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("C,Cout,res,relu", [(64, 64, True, False), (32, 64, False, True), (64, 32, True, True)])
def test_conv_instancenorm_autograd_vs_torch(C, Cout, res, relu):
    """Training block of the SemSegE2VID trunk (conv -> InstanceNorm -> (+x) -> ReLU) forward AND backward on the
    hand-written kernels (tcgen05 conv / dgrad / wgrad + InstanceNorm Jacobian) against torch autograd in float64."""
    import torch.nn.functional as F
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C + 2 * Cout)
    B, H, W = 2, 12, 20
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(Cout, C, 3, 3, generator=g, dtype=torch.float64) / (9 * C) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True)
    r = torch.randn(B, Cout, H, W, generator=g, dtype=torch.float64, requires_grad=True) if res else None
    xg = x.detach().float().cuda().requires_grad_(True)
    wg = w.detach().float().cuda().requires_grad_(True)
    bg = b.detach().float().cuda().requires_grad_(True)
    rg = r.detach().float().cuda().requires_grad_(True) if res else None
    yg = ops.conv_in_autograd(xg, wg, bg, rg, padding=1, relu=relu)
    y = F.instance_norm(F.conv2d(x, w, b, padding=1), eps=1e-5)
    if res:
        y = y + r
    if relu:
        # ReLU is discontinuous: a pre-activation within TF32 rounding of zero may land on the other side.  The reference
        # uses the kernel's own active set (the two forwards differ by < 1e-2 there), which makes the gradients comparable.
        y = y * (yg.detach().cpu() > 0).double()
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    yg.backward(dy.float().cuda())
    assert float((yg.detach().cpu().double() - y.detach()).abs().max()) < 1e-2
    for got, ref, name in ((xg.grad, x.grad, "dx"), (wg.grad, w.grad, "dW")):
        err = float((got.cpu().double() - ref).abs().max())
        assert err < 1e-2 * float(ref.abs().max()) + 1e-4, (name, err, float(ref.abs().max()))
    if res:
        assert float((rg.grad.cpu().double() - r.grad).abs().max()) < 1e-6
    assert float(bg.grad.abs().max()) < 1e-2 * float(w.grad.abs().max())      # mathematically zero


@pytest.mark.gpu
@pytest.mark.parametrize("C1,C2", [(64, 128), (32, 64), (16, 0)])
def test_upsample2x_cat_matches_torch_forward_and_backward(C1, C2):
    """oess_upsample2x_cat_nhwc (+ _bwd) against f.interpolate(nearest, x2) + torch.cat (style_networks.py:148-158): exact."""
    import torch.nn.functional as f
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C1 + C2)
    x = torch.randn(2, C1, 7, 9, generator=g).cuda().requires_grad_(True)
    s = torch.randn(2, C2, 14, 18, generator=g).cuda().requires_grad_(True) if C2 else None
    out = ops.upsample2x_cat(x, s)
    up = f.interpolate(x, scale_factor=2, mode='nearest')
    ref = torch.cat([up, s], dim=1) if C2 else up
    assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, ref)
    w = torch.randn(ref.shape, generator=g).cuda()
    gx, *gs = torch.autograd.grad((out * w).sum(), [x] + ([s] if C2 else []))
    rx, *rs = torch.autograd.grad((ref * w).sum(), [x] + ([s] if C2 else []))
    torch.testing.assert_close(gx, rx, rtol=1e-6, atol=1e-6)
    if C2:
        assert torch.equal(gs[0], rs[0])
