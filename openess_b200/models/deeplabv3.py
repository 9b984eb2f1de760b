"""DeepLabv3-ResNet-50: mirror of models/deeplabv3.py:86-189 (`DeepLabHead`, `deeplabv3_resnet50`),
:295-348 (`ASPPConv`, `ASPPPooling`, `ASPP`) and models/_resnet.py (ResNet-50 v1.5 = torchvision's) with the SAME module
tree and state_dict keys (`backbone.*` through layer4, `classifier.ASPP.*`, `classifier.pixel_feature.*`,
`classifier.classifier.*`, `classifier.text_embeddings`, `linear_probe.*`), so `torch.load(pretrained_backbone)['model_recon']`
loads with strict=True (deeplabv3.py:158-160).  Reference quirks kept: any `output_stride` != 8 (the YAMLs say 32) gives
replace_stride_with_dilation=[False, False, True] = output stride 16 (:137-142); `pixel_feature` is never used in forward.

What changes on the B200:
  * FROZEN backbone (`if_finetuning and frozen_backbone`, `if_linear_probing`: BASELINE config 4 / linear probing; and every
    no-grad call such as `val_step`): conv1..layer4 run through models/_tc_resnet.py -- 49 of its 53 convs as tcgen05
    implicit GEMMs (strided 3x3 / 1x1 via element-strided TMA boxes), BatchNorm in the module's mode (batch statistics +
    running-stat updates in train mode, folded in eval mode);
  * eval + no-grad (validation, test.py): the head also runs on the tensor cores -- ASPP branches, projection and the
    3x3 classifier conv with folded BN + ReLU epilogues;
  * the trainable head under autograd (fine-tuning, BASELINE config 4): every conv -> BatchNorm -> ReLU block is
    `ops.conv_bn_autograd` (tcgen05 forward / backward-data / backward-weight + BatchNorm Jacobian kernels); the one-pixel
    image-pooling branch, Dropout and the K-class text-embedding conv stay torch ops.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models._utils import IntermediateLayerGetter
from torchvision.models.resnet import Bottleneck, ResNet

from . import _tc_resnet as _tcr

USE_TENSOR_CORES = os.environ.get("OESS_DEEPLAB_TC", "1") != "0"
TRAIN_ON_TENSOR_CORES = os.environ.get("OESS_DEEPLAB_TC_TRAIN", "1") != "0"     # trainable head (fwd + bwd) on own kernels


class ASPPConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, dilation):
        super().__init__(nn.Conv2d(in_channels, out_channels, 3, padding=dilation, dilation=dilation, bias=False),
                         nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class ASPPPooling(nn.Sequential):
    def __init__(self, in_channels, out_channels):
        super().__init__(nn.AdaptiveAvgPool2d(1), nn.Conv2d(in_channels, out_channels, 1, bias=False),
                         nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))

    def forward(self, x):
        size = x.shape[-2:]
        x = super().forward(x)
        return F.interpolate(x, size=size, mode='bilinear', align_corners=False)


class ASPP(nn.Module):
    def __init__(self, in_channels, atrous_rates):
        super().__init__()
        out_channels = 256
        modules = [nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels),
                                 nn.ReLU(inplace=True))]
        for rate in tuple(atrous_rates):
            modules.append(ASPPConv(in_channels, out_channels, rate))
        modules.append(ASPPPooling(in_channels, out_channels))
        self.convs = nn.ModuleList(modules)
        self.project = nn.Sequential(nn.Conv2d(5 * out_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels),
                                     nn.ReLU(inplace=True), nn.Dropout(0.1))

    def forward(self, x):
        return self.project(torch.cat([conv(x) for conv in self.convs], dim=1))


class DeepLabHead(nn.Module):
    def __init__(self, text_embeddings_path, text_categories, in_channels, num_classes, aspp_dilate=[12, 24, 36]):
        super().__init__()
        self.ASPP = ASPP(in_channels, aspp_dilate)
        self.pixel_feature = nn.Conv2d(256, 512, 3, padding=1, bias=False)          # never used in forward (deeplabv3.py:94)
        self.classifier = nn.Sequential(nn.Conv2d(256, 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(inplace=True))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.text_embeddings_path = text_embeddings_path
        if text_embeddings_path is None:
            self.text_embeddings = nn.Parameter(torch.zeros(text_categories, 512))
            nn.init.normal_(self.text_embeddings, mean=0.0, std=0.01)
        else:
            self.register_buffer('text_embeddings', torch.randn(text_categories, 512))
            loaded = torch.load(text_embeddings_path, map_location='cpu')
            self.text_embeddings[:, :] = loaded[:, :]
        self._cache = _tcr.PackedConvCache()

    def forward(self, feature):
        x = feature['out']
        if (USE_TENSOR_CORES and x.is_cuda and not torch.is_grad_enabled() and not self.training
                and x.dtype == torch.float32):
            feature = self._aspp_tc(x)
            cls = _tcr.conv_bn(self._cache, feature, self.classifier[0], self.classifier[1], True)
        elif (USE_TENSOR_CORES and TRAIN_ON_TENSOR_CORES and x.is_cuda and torch.is_grad_enabled() and self.training
              and x.dtype == torch.float32):
            feature, cls = self._train_tc(x)
        else:
            feature = self.ASPP(x)
            cls = self.classifier(feature)
        logits = F.conv2d(cls, self.text_embeddings[:, :, None, None])
        return logits, feature

    def _train_tc(self, x):
        """Training forward (fine-tuning / pretraining of the head) with every conv -> BatchNorm -> ReLU block as
        `ops.conv_bn_autograd`: tcgen05 forward / backward-data / backward-weight convolutions + BatchNorm Jacobian kernels.
        The image-pooling branch (one pixel), Dropout and the K-class text-embedding conv stay torch ops."""
        from .. import ops as _ops
        a = self.ASPP
        cl = torch.channels_last
        x = x.contiguous(memory_format=cl)
        res = [_ops.conv_bn_autograd(x, a.convs[i][0], a.convs[i][1], relu=True) for i in range(4)]
        res.append(a.convs[4](x))
        cat = torch.cat(res, dim=1).contiguous(memory_format=cl)
        feature = a.project[3](_ops.conv_bn_autograd(cat, a.project[0], a.project[1], relu=True))      # Dropout(0.1)
        cls = _ops.conv_bn_autograd(feature, self.classifier[0], self.classifier[1], relu=True)
        return feature, cls

    def _aspp_tc(self, x):
        """ASPP.forward in eval mode on the tensor cores (folded BN + ReLU epilogues; Dropout is the identity in eval)."""
        a = self.ASPP
        x = x.contiguous(memory_format=torch.channels_last)
        res = [_tcr.conv_bn(self._cache, x, a.convs[i][0], a.convs[i][1], True) for i in range(4)]
        res.append(a.convs[4](x))                                                   # image pooling branch: 1 pixel, torch
        cat = torch.cat(res, dim=1).contiguous(memory_format=torch.channels_last)
        return _tcr.conv_bn(self._cache, cat, a.project[0], a.project[1], True)


class deeplabv3_resnet50(nn.Module):
    def __init__(self, num_classes, text_embeddings_path, output_stride, pretrained_backbone, if_linear_probing=False,
                 if_finetuning=False, frozen_backbone=False):
        super().__init__()
        if output_stride == 8:
            replace_stride_with_dilation, aspp_dilate = [False, True, True], [12, 24, 36]
        else:
            replace_stride_with_dilation, aspp_dilate = [False, False, True], [6, 12, 18]
        backbone = ResNet(Bottleneck, [3, 4, 6, 3], replace_stride_with_dilation=replace_stride_with_dilation)
        self.backbone = IntermediateLayerGetter(backbone, return_layers={'layer4': 'out'})
        self.classifier = DeepLabHead(text_embeddings_path, num_classes, 2048, num_classes, aspp_dilate)
        if pretrained_backbone != '':
            pretrained = torch.load(pretrained_backbone, map_location='cpu')
            self.load_state_dict(pretrained['model_recon'], strict=True)
        self.if_linear_probing = if_linear_probing
        if self.if_linear_probing:
            for param in self.backbone.parameters():
                param.requires_grad = False
            for param in self.classifier.parameters():
                param.requires_grad = False
            self.linear_probe = nn.Conv2d(num_classes, num_classes, 1)
        self.if_finetuning = if_finetuning
        if self.if_finetuning and frozen_backbone:
            for param in self.backbone.parameters():
                param.requires_grad = False
            for param in self.classifier.parameters():
                param.requires_grad = True
        self._cache = _tcr.PackedConvCache()

    def _backbone(self, x):
        tc_ok = (USE_TENSOR_CORES and x.is_cuda and not x.requires_grad and x.dtype == torch.float32
                 and (_tcr.frozen(self.backbone) or not torch.is_grad_enabled()))
        if tc_ok:
            with torch.no_grad():
                return {'out': _tcr.resnet_stages(self._cache, self.backbone, x)}
        return self.backbone(x)

    def forward(self, x):
        input_shape = x.shape[-2:]
        features = self._backbone(x)                                                # [B, 2048, H/16, W/16]
        logist, feats = self.classifier(features)
        if USE_TENSOR_CORES and logist.is_cuda and logist.dtype == torch.float32 and feats.dtype == torch.float32:
            from .. import ops as _ops                                              # own resize: gather backward, no atomics
            logist = _ops.bilinear_resize(logist, input_shape)
            feats = _ops.bilinear_resize(feats, input_shape)
        else:
            logist = F.interpolate(logist, size=input_shape, mode='bilinear', align_corners=False)
            feats = F.interpolate(feats, size=input_shape, mode='bilinear', align_corners=False)
        if self.if_linear_probing:
            logist = self.linear_probe(logist)
        return logist, feats
