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

    def get(self, conv, bn):
        fold = bn is not None and not bn.training
        key = (id(conv), fold)
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
            hit = (ver, _tc.conv2d_pack(w), b)
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
        return _tc.conv2d_tc(x, wp, b, k, s, p, d, relu=relu, residual=residual)
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


def resnet_stages(cache, net, x):
    """conv1 / bn1 / relu / maxpool as torch ops (Cin = 3), layer1..4 on the tensor cores; returns layer4 channels-last."""
    x = net.relu(net.bn1(net.conv1(x)))
    x = net.maxpool(x).contiguous(memory_format=torch.channels_last)
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        for blk in layer:
            x = bottleneck(cache, blk, x)
    return x


def frozen(module):
    return not any(p.requires_grad for p in module.parameters())
