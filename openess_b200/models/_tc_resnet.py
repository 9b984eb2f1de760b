"""Tensor-core execution of frozen torchvision-style ResNet stages (Bottleneck blocks), shared by the dilated ResNet-50
teacher (models/image_model.py, row a13) and the frozen DeepLabv3 backbone (models/deeplabv3.py, row a12).

Every conv runs on `oess_conv2d_nhwc_tf32` (tcgen05 implicit GEMM, channels-last).  BatchNorm follows the module's mode:
train mode = batch statistics + running-stat update through `oess_batchnorm_nhwc` (what the OpenESS trainers do to frozen
networks: `.train()` is called on every model each step, pretrain_trainer.py:370-371), eval mode = folded into the conv
weights with bias / residual / ReLU in the conv epilogue."""
import torch

from .. import ops as _tc


FUSE_BN_STATS = True      # batch statistics accumulated in the conv epilogue instead of a separate pass


class PackedConvCache:
    """Packed (and, in eval mode, BN-folded) weights per (conv, mode); rebuilt when parameters change or move."""

    def __init__(self):
        self._packed = {}

    def get(self, conv, bn, pad_cin=0, bf16=False):
        """pad_cin: zero-pad the input channels to this count (the 3-channel stem runs on an 8-channel repack).
        bf16: weights packed as bfloat16 for `oess_conv2d_nhwc_bf16`."""
        fold = bn is not None and not bn.training
        key = (id(conv), fold, bf16)
        ver = (conv.weight.data_ptr(), conv.weight._version, conv.weight.device,
               (bn.running_var._version, bn.weight._version, bn.bias._version) if fold else None)
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            w = conv.weight.detach()
            b = None if conv.bias is None else conv.bias.detach().float().contiguous()
            if fold:
                scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
                w = w * scale[:, None, None, None]
                b0 = 0 if b is None else b
                b = ((b0 - bn.running_mean) * scale + bn.bias.detach()).float().contiguous()
            if pad_cin > w.shape[1]:
                wp = torch.zeros(w.shape[0], pad_cin, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
                wp[:, :w.shape[1]] = w
                w = wp
            hit = (ver, _tc.conv2d_pack_bf16(w) if bf16 else _tc.conv2d_pack(w), b)
            self._packed[key] = hit
        return hit[1], hit[2]


def conv_supported(conv):
    return (conv.groups == 1 and conv.padding_mode == 'zeros' and conv.in_channels % 4 == 0 and conv.in_channels >= 16
            and conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1]
            and conv.padding[0] == conv.padding[1] and conv.dilation[0] == conv.dilation[1])


def conv_bn(cache, x, conv, bn, relu, residual=None):
    """act(bn(conv(x)) + residual) on a channels-last CUDA tensor, no autograd."""
    wp, b = cache.get(conv, bn)
    k, s, p, d = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
    if bn is None or not bn.training:
        # eval mode: nothing re-normalises between the layers, so the activation is stored TF32-rounded (round-to-nearest)
        # instead of being truncated by the next conv's tensor-core read (a -2^-12 relative bias per layer otherwise)
        return _tc.conv2d_tc(x, wp, b, k, s, p, d, relu=relu, residual=residual, round_out=True)
    if FUSE_BN_STATS and bn.weight.numel() % 4 == 0:
        return _tc.conv_bn_train(x, wp, b, k, s, p, d, bn, residual=residual, relu=relu)
    y = _tc.conv2d_tc(x, wp, b, k, s, p, d)
    return _tc.batchnorm_nhwc_(y, bn, residual=residual, relu=relu)


def bottleneck(cache, blk, x):
    """torchvision.models.resnet.Bottleneck.forward / models/_resnet.py:94-114."""
    out = conv_bn(cache, x, blk.conv1, blk.bn1, True)
    out = conv_bn(cache, out, blk.conv2, blk.bn2, True)
    identity = x
    if blk.downsample is not None:
        identity = conv_bn(cache, x, blk.downsample[0], blk.downsample[1], False)
    return conv_bn(cache, out, blk.conv3, blk.bn3, True, residual=identity)


def conv_bn_bf16(cache, x_bf, conv, bn, relu, residual=None, want_f32=True):
    """conv_bn with bfloat16 conv operands; returns (y fp32 or None, y bf16)."""
    wp, b = cache.get(conv, bn, bf16=True)
    k, s, p, d = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
    if bn is None or not bn.training:
        return _tc.conv2d_tc_bf16(x_bf, wp, b, k, s, p, d, relu=relu, residual=residual, want_f32=want_f32)
    return _tc.conv_bn_train_bf16(x_bf, wp, b, k, s, p, d, bn, residual=residual, relu=relu, want_f32=want_f32)


def bottleneck_bf16(cache, blk, x, x_bf):
    """`bottleneck` with bf16 operands: activations inside the block exist only as bf16, the block output (the next block's
    identity) as fp32 + bf16; every sum (accumulators, BatchNorm statistics, residual add) is fp32."""
    _, out = conv_bn_bf16(cache, x_bf, blk.conv1, blk.bn1, True, want_f32=False)
    _, out = conv_bn_bf16(cache, out, blk.conv2, blk.bn2, True, want_f32=False)
    identity = x
    if blk.downsample is not None:
        identity, _ = conv_bn_bf16(cache, x_bf, blk.downsample[0], blk.downsample[1], False)
    return conv_bn_bf16(cache, out, blk.conv3, blk.bn3, True, residual=identity)


def bf16_ok(net):
    """every block a Bottleneck whose convs the bf16 kernel takes and whose BatchNorms have C % 4 == 0"""
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            if not hasattr(blk, "conv3"):
                return False
            convs = [blk.conv1, blk.conv2, blk.conv3] + ([blk.downsample[0]] if blk.downsample is not None else [])
            if not all(conv_supported(c) and c.in_channels % 8 == 0 and c.out_channels % 4 == 0 for c in convs):
                return False
    return True


def basic_block(cache, blk, x):
    """torchvision.models.resnet.BasicBlock.forward / models/_resnet.py:55-71 (ResNet-18 / 34)."""
    out = conv_bn(cache, x, blk.conv1, blk.bn1, True)
    identity = x
    if blk.downsample is not None:
        identity = conv_bn(cache, x, blk.downsample[0], blk.downsample[1], False)
    return conv_bn(cache, out, blk.conv2, blk.bn2, True, residual=identity)


STEM_TC = True            # False: conv1 / bn1 / relu / maxpool as torch ops (the round-1c formulation)


def stem(cache, net, x):
    """conv1 (7x7 stride 2, Cin = 3) + bn1 + relu + maxpool (models/_resnet.py:134-137, 199-202) -> channels-last.
    Cin = 3 is too thin for a 16-byte TMA pixel: the planes are repacked to 8 zero-padded channels-last channels
    (`oess_planes_to_nhwc_padded`) and run through the same tcgen05 conv kernel (BN folded or batch statistics in the
    epilogue, ReLU fused); the max pool is `oess_maxpool3x3s2_nhwc`."""
    conv, bn = net.conv1, net.bn1
    if not (STEM_TC and conv.in_channels <= 8 and conv.groups == 1 and conv.kernel_size[0] == conv.kernel_size[1]
            and conv.kernel_size[0] ** 2 <= 64 and isinstance(bn, torch.nn.BatchNorm2d)
            and type(net.maxpool) is torch.nn.MaxPool2d and net.maxpool.kernel_size == 3 and net.maxpool.stride == 2
            and net.maxpool.padding == 1 and net.maxpool.dilation == 1 and not net.maxpool.ceil_mode):
        x = net.relu(net.bn1(net.conv1(x)))
        return net.maxpool(x).contiguous(memory_format=torch.channels_last)
    wp, b = cache.get(conv, bn, pad_cin=8)
    x8 = _tc.planes_to_nhwc_padded(x, 8)
    k, s, p, d = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
    if not bn.training:
        y = _tc.conv2d_tc(x8, wp, b, k, s, p, d, relu=True, round_out=True)
    elif FUSE_BN_STATS:
        y = _tc.conv_bn_train(x8, wp, b, k, s, p, d, bn, relu=True)
    else:
        y = _tc.batchnorm_nhwc_(_tc.conv2d_tc(x8, wp, b, k, s, p, d), bn, relu=True)
    return _tc.maxpool3x3s2_nhwc(y)


def resnet_stages(cache, net, x, bf16=False):
    """Stem + layer1..4 on hand-written kernels (tcgen05 convs, fused BN, own max pool); returns layer4 channels-last.
    bf16 (frozen Bottleneck networks): bfloat16 conv operands after the stem, fp32 accumulation / statistics / residuals."""
    x = stem(cache, net, x)
    if bf16 and bf16_ok(net):
        x_bf = x.to(torch.bfloat16)
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                x, x_bf = bottleneck_bf16(cache, blk, x, x_bf)
        return x
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            x = bottleneck(cache, blk, x) if hasattr(blk, "conv3") else basic_block(cache, blk, x)
    return x


def frozen(module):
    return not any(p.requires_grad for p in module.parameters())
