"""tcgen05 implicit-GEMM convolution (oess_conv2d_nhwc_tf32) against torch's conv2d evaluated in float64 on the CPU.
Tolerance: TF32 operands, fp32 accumulate -> |err| <= 2e-3 * conv(|x|, |w|) (stated in include/openess_b200.h)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p,d,relu,res", [
    (1, 32, 64, 16, 32, 5, 2, 2, 1, True, False),     # E2VID encoder 1 (unet.py:128-135): 5x5 stride 2 + ReLU
    (2, 64, 128, 22, 40, 5, 2, 2, 1, True, False),    # encoder 2, ragged tiles
    (1, 128, 256, 11, 20, 5, 2, 2, 1, True, False),   # encoder 3, odd input size
    (1, 64, 64, 9, 17, 3, 1, 1, 1, False, True),      # 3x3 + residual (ResidualBlock, submodules.py:140-172)
    (1, 64, 64, 12, 20, 3, 1, 2, 2, True, False),     # dilated 3x3 (ResNet layer with replace_stride_with_dilation)
    (1, 256, 64, 10, 16, 1, 1, 0, 1, True, False),    # 1x1 bottleneck reduce
    (1, 20, 24, 8, 16, 3, 1, 1, 1, False, False),     # Cin not a multiple of 32 (zero-filled chunk), Cout % 16 != 0
    (1, 8, 10, 8, 16, 3, 2, 1, 1, True, False),       # scalar store path (Cout % 4 != 0)
    (3, 128, 320, 40, 72, 3, 1, 1, 1, True, True),    # CTA pairs (cta_group::2; BN = 256, K = 1152): 25 pixel tiles (odd: a dummy peer tile), ragged second Cout tile
    (6, 32, 96, 40, 72, 3, 1, 1, 1, False, True),     # many tiles per CTA, BN = 128, two K blocks per stage with an odd tail (9 blocks), ragged Cout
])
def test_conv2d_tc(B, Cin, Cout, H, W, k, s, p, d, relu, res):
    from openess_b200 import ops
    g = torch.Generator().manual_seed(Cin * 7 + k)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p, dilation=d)
    bound = 2e-3 * F.conv2d(x.abs().double(), w.abs().double(), None, stride=s, padding=p, dilation=d) + 1e-6
    r = None
    if res:
        r = torch.randn(ref.shape, generator=g)
        ref = ref + r.double()
    if relu:
        ref = ref.clamp_min(0)
    y = ops.conv2d_tc(x.cuda(), ops.conv2d_pack(w.cuda()), b.cuda(), k, s, p, d, relu, None if r is None else r.cuda())
    assert tuple(y.shape) == tuple(ref.shape)
    err = (y.cpu().double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {float(err.max())}, bound min {float(bound.min())}"
    assert float(err.max()) < 0.05 * float(ref.abs().max())


def test_conv2d_tc_exact_on_tf32_representable_inputs():
    from openess_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randint(-3, 4, (2, 32, 13, 19), generator=g).float()
    w = torch.randint(-2, 3, (48, 32, 5, 5), generator=g).float() / 8.0
    b = torch.randint(-4, 5, (48,), generator=g).float() / 2.0
    ref = F.conv2d(x, w, b, stride=2, padding=2).clamp_min(0)
    y = ops.conv2d_tc(x.cuda(), ops.conv2d_pack(w.cuda()), b.cuda(), 5, 2, 2, 1, True)
    assert torch.equal(y.cpu(), ref)


@pytest.mark.parametrize("C,Cout,H,W,k,s,p,d", [(64, 256, 22, 37, 1, 1, 0, 1), (64, 64, 19, 30, 3, 1, 2, 2), (128, 32, 17, 20, 3, 2, 1, 1),
                                                 (128, 256, 104, 136, 3, 1, 1, 1)])     # CTA pairs: 2 x 59 pair tiles (117 pixel tiles per sample: a dummy peer tile), K = 1152
def test_conv_bn_train_fused_stats_vs_torch(C, Cout, H, W, k, s, p, d):
    """conv + train-mode BatchNorm with the batch statistics accumulated in the conv's TMEM epilogue, against torch
    (fp64 conv on the CPU, then torch.nn.BatchNorm2d in train mode)."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C + Cout)
    x = torch.randn(2, C, H, W, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5
    bn_ref = torch.nn.BatchNorm2d(Cout).double()
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.normal_(0, 0.3)
    bn = torch.nn.BatchNorm2d(Cout)
    bn.load_state_dict({k_: v.float() if v.is_floating_point() else v for k_, v in bn_ref.state_dict().items()})
    bn = bn.cuda().train()
    bn_ref.train()
    r = torch.randn(2, Cout, (H + 2 * p - d * (k - 1) - 1) // s + 1, (W + 2 * p - d * (k - 1) - 1) // s + 1, generator=g)
    with torch.no_grad():
        ref = (bn_ref(F.conv2d(x.double(), w.double(), None, stride=s, padding=p, dilation=d)) + r.double()).relu()
    y = ops.conv_bn_train(x.cuda(), ops.conv2d_pack(w.cuda()), None, k, s, p, d, bn, residual=r.cuda(), relu=True)
    assert float((y.cpu().double() - ref).abs().max()) < 1e-2                      # TF32 conv, normalised to unit scale
    torch.testing.assert_close(bn.running_mean.cpu().double(), bn_ref.running_mean, atol=2e-4, rtol=1e-3)
    torch.testing.assert_close(bn.running_var.cpu().double(), bn_ref.running_var, atol=2e-4, rtol=2e-3)
    assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("C,Cout,k,p,d", [(32, 64, 3, 1, 1), (64, 32, 3, 2, 2), (128, 64, 1, 0, 1), (32, 32, 5, 2, 1)])
def test_conv2d_dgrad_is_the_forward_kernel_on_repacked_weights(C, Cout, k, p, d):
    """dL/dx of a stride-1 conv == conv2d_tc(dL/dy, rot180(W)^T, padding = d (k - 1) - p): the backward-data pass needs no
    new kernel.  Checked against torch autograd in float64."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C + Cout + k)
    x = torch.randn(2, C, 14, 19, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, C, k, k, generator=g, dtype=torch.float64) / (C * k * k) ** 0.5
    y = F.conv2d(x, w, None, padding=p, dilation=d)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    wp, pad = ops.conv2d_dgrad_pack(w.float().cuda(), p, d)
    dx = ops.conv2d_tc(dy.float().cuda(), wp, None, k, 1, pad, d)
    assert tuple(dx.shape) == tuple(x.shape)
    bound = 2e-3 * F.conv2d(dy.abs(), w.abs().flip(2, 3).permute(1, 0, 2, 3), None, padding=pad, dilation=d) + 1e-6
    assert bool(((dx.cpu().double() - x.grad).abs() <= bound).all())


@pytest.mark.parametrize("B,C,Cout,H,W,k,p,d", [
    (2, 64, 32, 12, 20, 3, 1, 1),          # SemSegE2VID-like 3x3
    (1, 2048, 256, 11, 16, 1, 0, 1),       # teacher decoder 1x1 conv 2048 -> 256 (image_model.py:121-124)
    (2, 32, 48, 9, 36, 3, 2, 2),           # dilated, Cout not a multiple of 128, W not a multiple of 32
    (1, 16, 16, 8, 8, 5, 2, 1),            # 5x5
    (3, 256, 512, 7, 12, 3, 1, 1),         # DeepLab classifier 3x3 256 -> 512: several co tiles, split K over images
])
def test_conv2d_autograd_tensor_cores_vs_torch(B, C, Cout, H, W, k, p, d):
    """forward + backward-data + backward-weight of ops.conv2d_tc_autograd against torch autograd in float64.
    Tolerance: TF32 operands -> 2e-3 * (sum of |terms|), evaluated with the same convolutions on absolute values."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(B * C + Cout + k)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(Cout, C, k, k, generator=g, dtype=torch.float64) / (C * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, b, padding=p, dilation=d)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    xg = x.detach().float().cuda().requires_grad_(True)
    wg = w.detach().float().cuda().requires_grad_(True)
    bg = b.detach().float().cuda().requires_grad_(True)
    yg = ops.conv2d_tc_autograd(xg, wg, bg, p, d)
    yg.backward(dy.float().cuda())
    xa, wa, dya = x.detach().abs(), w.detach().abs(), dy.abs()
    bound_w = 2e-3 * torch.autograd.grad(F.conv2d(xa, wa.requires_grad_(True), None, padding=p, dilation=d), wa, dya)[0] + 1e-5
    bound_x = 2e-3 * torch.autograd.grad(F.conv2d(xa.requires_grad_(True), wa.detach(), None, padding=p, dilation=d), xa, dya)[0] + 1e-5
    assert bool(((wg.grad.cpu().double() - w.grad).abs() <= bound_w).all()), float((wg.grad.cpu().double() - w.grad).abs().max())
    assert bool(((xg.grad.cpu().double() - x.grad).abs() <= bound_x).all())
    torch.testing.assert_close(bg.grad.cpu().double(), b.grad, atol=1e-4, rtol=1e-5)
    assert float((yg.detach().cpu().double() - y.detach()).abs().max()) < 1e-2


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p,d,relu,res", [
    (1, 64, 128, 22, 40, 5, 2, 2, 1, True, False),    # E2VID encoder 2 with bf16 operands
    (2, 128, 256, 11, 20, 5, 2, 2, 1, True, False),   # encoder 3, two 64-channel chunks, odd input size
    (1, 64, 64, 12, 20, 3, 1, 2, 2, True, False),     # dilated 3x3 (teacher layer3 / layer4)
    (1, 256, 64, 10, 16, 1, 1, 0, 1, True, False),    # 1x1 bottleneck reduce, four chunks
    (1, 64, 256, 9, 17, 1, 1, 0, 1, True, True),      # 1x1 expand + fp32 residual
    (1, 32, 64, 16, 32, 5, 2, 2, 1, True, False),     # Cin = 32: half of the 64-element K block is TMA zero fill
    (1, 24, 20, 8, 16, 3, 1, 1, 1, False, False),     # Cin % 64 != 0, Cout % 16 != 0
    (3, 128, 320, 40, 72, 3, 1, 1, 1, True, True),    # CTA pairs with bf16 operands, ragged second Cout tile, dummy peer tile
])
def test_conv2d_tc_bf16_operands(B, Cin, Cout, H, W, k, s, p, d, relu, res):
    """oess_conv2d_nhwc_bf16 against torch's conv2d in float64 on the bf16-ROUNDED operands (products of bf16 values are exact
    in fp32: only the accumulation order differs -> 1e-5 * conv(|x|, |w|)); the bf16 copy of the result within one bf16 ulp."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(Cin * 5 + k)
    x = torch.randn(B, Cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p, dilation=d)
    bound = 1e-5 * F.conv2d(x.abs().double(), w.abs().double(), None, stride=s, padding=p, dilation=d) + 1e-6
    r = None
    if res:
        r = torch.randn(ref.shape, generator=g)
        ref = ref + r.double()
    if relu:
        ref = ref.clamp_min(0)
    xb = x.cuda().contiguous(memory_format=torch.channels_last)
    y, yb = ops.conv2d_tc_bf16(xb, ops.conv2d_pack_bf16(w.cuda().float()), b.cuda(), k, s, p, d, relu,
                               None if r is None else r.cuda())
    assert tuple(y.shape) == tuple(ref.shape) and yb.dtype == torch.bfloat16
    err = (y.cpu().double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {float(err.max())}, bound min {float(bound.min())}"
    assert torch.equal(yb, y.to(torch.bfloat16))
    only_bf = ops.conv2d_tc_bf16(xb, ops.conv2d_pack_bf16(w.cuda().float()), b.cuda(), k, s, p, d, relu,
                                 None if r is None else r.cuda(), want_f32=False)
    assert only_bf[0] is None and torch.equal(only_bf[1], yb)


def test_conv_bn_train_bf16_vs_torch():
    """bf16-operand conv + train-mode BatchNorm (statistics from the fp32 accumulators) against torch on the rounded operands."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(11)
    C, Cout, H, W, k = 64, 256, 22, 37, 1
    x = torch.randn(2, C, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, C, k, k, generator=g) / C ** 0.5).to(torch.bfloat16)
    res = torch.randn(2, Cout, H, W, generator=g)
    bn_ref = torch.nn.BatchNorm2d(Cout).double()
    bn = torch.nn.BatchNorm2d(Cout).cuda()
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
        bn_ref.bias.copy_(torch.randn(Cout, generator=g) * 0.1)
        bn.weight.copy_(bn_ref.weight.float())
        bn.bias.copy_(bn_ref.bias.float())
        ref = (bn_ref(F.conv2d(x.double(), w.double())) + res.double()).clamp_min(0)
    xb = x.cuda().contiguous(memory_format=torch.channels_last)
    y, yb = ops.conv_bn_train_bf16(xb, ops.conv2d_pack_bf16(w.cuda().float()), None, k, 1, 0, 1, bn, residual=res.cuda(), relu=True)
    assert float((y.cpu().double() - ref).abs().max()) < 2e-4
    assert torch.equal(yb, y.to(torch.bfloat16))
    torch.testing.assert_close(bn.running_var.cpu().double(), bn_ref.running_var, rtol=1e-4, atol=1e-5)
    y2, yb2 = ops.conv_bn_train_bf16(xb, ops.conv2d_pack_bf16(w.cuda().float()), None, k, 1, 0, 1, bn, residual=res.cuda(),
                                     relu=True, want_f32=False)
    assert y2 is None and torch.equal(yb2, yb)


@pytest.mark.parametrize("B,C,Cout,H,W,k", [(2, 5, 32, 21, 37, 5), (1, 3, 64, 16, 48, 3), (1, 8, 32, 9, 16, 7)])
def test_conv2d_rowunfold_thin_input(B, C, Cout, H, W, k):
    """Thin-input 'same' convolution with the taps of a kernel row folded into the channel dimension (E2VID head, unet.py:126-127)
    against torch's conv2d in float64."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(C * 11 + k)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2).clamp_min(0)
    bound = 2e-3 * F.conv2d(x.abs().double(), w.abs().double(), None, padding=k // 2) + 1e-6
    wpad = torch.zeros(Cout, 8, k, k)
    wpad[:, :C] = w
    x8 = ops.planes_to_nhwc_padded_w(x.cuda(), 8, k // 2)
    assert tuple(x8.shape) == (B, H, W + k - 1, 8)
    y = ops.conv2d_rowunfold(x8, ops.conv2d_pack_rowunfold(wpad.cuda()), b.cuda(), k, k, W, relu=True)
    err = (y.cpu().double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {float(err.max())}"
