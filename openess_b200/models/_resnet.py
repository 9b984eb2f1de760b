"""ResNet-18 / 34 / 50 classifiers (SURVEY.md 8a row a14'): mirror of models/_resnet.py:117-260 (`ResNet`, `BasicBlock`
:34-71, `Bottleneck` :74-114, `resnet18` :225-234, `resnet34`, `resnet50`).  The reference file is torchvision's
resnet.py of 2019 with a string-valued `pretrained`; module tree and state_dict keys equal torchvision's, whose
`BasicBlock` / `Bottleneck` / `ResNet` classes are the parameter containers here (as in models/image_model.py and
models/deeplabv3.py of this package).  BASELINE config 2 names the ResNet-18 as the event encoder; no trainer of the
reference instantiates it, so it is built and measured module-level.

B200 forward (CUDA tensor, no autograd needed: grad disabled, or frozen parameters and an input that does not require
grad): every convolution -- the 7x7 stem through an 8-channel channels-last repack, the 3x3 / strided 3x3 / 1x1
downsample convs of the blocks -- is a tcgen05 implicit GEMM (`oess_conv2d_nhwc_tf32`), BatchNorm follows the module's
mode (train: batch statistics from the conv epilogue + running-stat update, eval: folded into the conv), max pool and
global average pool are `oess_maxpool3x3s2_nhwc` / `oess_global_avgpool_nhwc`, the classifier is `oess_gemm_tf32_ex`.
`forward_features` returns the layer4 map ([B, 512, H / 32, W / 32] for ResNet-18) for use as an encoder.
With autograd (trainable parameters) the torch formulation runs -- the reference trains no such network.

`pretrained` follows models/_resnet.py:212-222: '' -> random init, 'imagenet' -> download (no network here: raises),
any other string -> `torch.load(path)` + `load_state_dict(strict=False)`.  Deviation, documented: the reference's
`resnet18(pretrained=False)` default reaches `torch.load(False)` and crashes; False / None mean '' here."""
import torch
import torch.nn as nn
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet as _TVResNet

from .. import ops as _tc
from . import _tc_resnet as _tcr

__all__ = ['ResNet', 'BasicBlock', 'Bottleneck', 'resnet18', 'resnet34', 'resnet50']


class ResNet(_TVResNet):
    """models/_resnet.py:117-209."""

    def __init__(self, block, layers, num_classes=1000, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None):
        super().__init__(block, layers, num_classes=num_classes, zero_init_residual=zero_init_residual, groups=groups,
                         width_per_group=width_per_group, replace_stride_with_dilation=replace_stride_with_dilation,
                         norm_layer=norm_layer)
        self._cache = _tcr.PackedConvCache()

    def _tc_ok(self, x):
        return (x.is_cuda and x.dtype == torch.float32 and self.groups == 1
                and (not torch.is_grad_enabled() or (not x.requires_grad and _tcr.frozen(self)))
                and all(isinstance(m, nn.BatchNorm2d) for m in (self.bn1, self.layer1[0].bn1)))

    def forward_torch(self, x, features_only=False):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if features_only:
            return x
        return self.fc(torch.flatten(self.avgpool(x), 1))

    def forward_features(self, x):
        """layer4 feature map (channels-last memory format on the tensor-core path)."""
        if self._tc_ok(x):
            with torch.no_grad():
                return _tcr.resnet_stages(self._cache, self, x)
        return self.forward_torch(x, features_only=True)

    def forward(self, x):
        if self._tc_ok(x):
            with torch.no_grad():
                f = _tcr.resnet_stages(self._cache, self, x)
                return _tc.gemm_tf32_ex(_tc.global_avgpool_nhwc(f), self.fc.weight, self.fc.bias)
        return self.forward_torch(x)


def _resnet(arch, block, layers, pretrained, progress, **kwargs):
    model = ResNet(block, layers, **kwargs)
    if pretrained == 'imagenet':
        raise RuntimeError(f"{arch}: pretrained='imagenet' downloads the torchvision weights (models/_resnet.py:214-215); "
                           "there is no network here -- pass the path of a local state_dict instead")
    if pretrained == '' or pretrained is False or pretrained is None:
        return model
    model.load_state_dict(torch.load(pretrained, map_location='cpu'), strict=False)
    return model


def resnet18(pretrained=False, progress=True, **kwargs):
    """models/_resnet.py:225-234."""
    return _resnet('resnet18', BasicBlock, [2, 2, 2, 2], pretrained, progress, **kwargs)


def resnet34(pretrained=False, progress=True, **kwargs):
    """models/_resnet.py:237-246."""
    return _resnet('resnet34', BasicBlock, [3, 4, 6, 3], pretrained, progress, **kwargs)


def resnet50(pretrained='', progress=True, **kwargs):
    """models/_resnet.py:249-258 (default 'imagenet' in the reference = a download; '' here)."""
    return _resnet('resnet50', Bottleneck, [3, 4, 6, 3], pretrained, progress, **kwargs)
