"""Dilated ResNet-50 teacher: mirror of models/image_model.py:90-143 (DilationFeatureExtractor) and
models/modules/resnet_encoder.py:8-38 (ResNetEncoder = torchvision ResNet-50 without fc / avgpool,
replace_stride_with_dilation=[True, True, True] -> output stride 4) with the SAME module tree and state_dict keys
(`encoder.*`, `decoder.0.*`), so the reference's teacher weights (dino / moco / swav / imagenet .pt files, adapted by
image_model.py:26-74) load with `load_state_dict(strict=True)`.

What changes on the B200 (`forward` on a CUDA tensor that does not require grad -- the encoder is frozen, every trainer
feeds it the raw frame):
  * all 52 bottleneck convolutions (1x1, dilated 3x3, downsample 1x1: 99 % of the 845 GFLOP / sample) run as tcgen05
    implicit GEMMs over channels-last activations (`oess_conv2d_nhwc_tf32`, TF32 operands / fp32 accumulate);
  * BatchNorm runs the way the trainers really run it -- `.train()` is called on the frozen teacher every step
    (pretrain_trainer.py:370-371), so BATCH statistics are used and the running statistics are updated -- as the fused
    channels-last kernels of `oess_batchnorm_nhwc` (stats -> finalize on device -> affine + residual + ReLU in place);
    in eval mode BN is folded into the conv and bias / residual / ReLU go into the conv epilogue (no BN pass at all);
  * the 3-channel 7x7 stem conv + maxpool (1 % of the FLOPs; Cin = 3 is too thin for a 128-byte TMA row) and the
    trainable decoder (1x1 conv 2048 -> 256 with autograd, x4 bilinear upsample, L2 normalise) stay torch ops.
Preprocessing is `None` in every trainer (pretrain_trainer.py:185-187): raw [0, 1] RGB goes in, as in the reference.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import Bottleneck, ResNet

from .. import ops as _tc
from . import _tc_resnet as _tcr

USE_TENSOR_CORES = os.environ.get("OESS_TEACHER_TC", "1") != "0"
# operand type of the frozen encoder's tensor-core convs after the stem: tf32 (default) or, with OESS_TEACHER_DTYPE=bf16, bfloat16
# (tcgen05 kind::f16; fp32 accumulation, BatchNorm statistics and residual adds).  Opt-in: measured 15 % faster in train mode
# (B = 8: 24.8 -> 21.0 ms; the 1 x 1 convs are bound by their fp32 output + BatchNorm passes, not by the MMA rate) for 6x the
# feature error of TF32 on the seeded network of tests/test_teacher.py (52 convs, each re-normalised by batch statistics).
TEACHER_BF16 = os.environ.get("OESS_TEACHER_DTYPE", "tf32") == "bf16"
TEACHER_GRAPH = os.environ.get("OESS_TEACHER_GRAPH", "1") != "0"


class ResNetEncoder(ResNet):
    """models/modules/resnet_encoder.py:8-38."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        del self.fc
        del self.avgpool
        self._cache = _tcr.PackedConvCache()

    def load_state_dict(self, state_dict, **kwargs):
        if "_graph_call" in self.__dict__:
            self.__dict__["_graph_call"].reset()
        state_dict.pop("fc.bias", None)
        state_dict.pop("fc.weight", None)
        return super().load_state_dict(state_dict, **kwargs)

    # ---- reference formulation (CPU / autograd / USE_TENSOR_CORES = False) ----
    def forward_torch(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.layer1(self.maxpool(x))
        return self.layer4(self.layer3(self.layer2(x)))

    # ---- tensor-core formulation (models/_tc_resnet.py) ----
    def forward_tc(self, x):
        return _tcr.resnet_stages(self._cache, self, x, bf16=TEACHER_BF16)

    def _graphed(self):
        """The frozen tensor-core forward (53 conv + BatchNorm blocks = ~210 launches issued from Python, ~11 ms of host time at
        any batch size) as a CUDA graph (training/graphs.py GraphedCall; OESS_TEACHER_GRAPH=0: eager).  Dropped whenever the
        module's mode, dtype switch, device or parameters change."""
        g = self.__dict__.get("_graph_call")
        if g is None:
            from ..training.graphs import GraphedCall

            def run(x):
                with torch.no_grad():
                    return self.forward_tc(x)
            g = GraphedCall(run, state=lambda: (self.training, TEACHER_BF16, self.conv1.weight._version,
                                                 self.layer4[-1].conv3.weight._version))
            self.__dict__["_graph_call"] = g
        return g

    def _apply(self, fn, *args, **kwargs):
        if "_graph_call" in self.__dict__:
            self.__dict__["_graph_call"].reset()
        return super()._apply(fn, *args, **kwargs)

    def forward(self, x):
        tc_ok = (USE_TENSOR_CORES and x.is_cuda and not x.requires_grad and x.dtype == torch.float32
                 and _tcr.frozen(self))
        if tc_ok:
            if TEACHER_GRAPH:
                return self._graphed()(x)
            with torch.no_grad():
                return self.forward_tc(x)
        return self.forward_torch(x)


def adapt_weights(architecture):
    """models/image_model.py:26-74 for weight files already on disk (`weights/<architecture>.pt`): same key rewriting per
    source (moco / swav / deepcluster / dino / pixpro / obow).  The reference additionally downloads a missing file with
    `requests` (:38-44): network ingest is out of scope -- a missing file raises FileNotFoundError here."""
    if architecture == "imagenet" or architecture is None:
        return None
    path = f"weights/{architecture}.pt"
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: place the teacher weights there (the reference downloads them at "
                                "run time, models/image_model.py:38-44)")
    weights = torch.load(path, map_location="cpu")
    if architecture == "obow":
        return weights["network"]
    if architecture == "pixpro":
        return {k.replace("module.encoder.", ""): v for k, v in weights["model"].items() if k.startswith("module.encoder.")}
    if architecture in ("moco_v1", "moco_v2", "moco_coco"):
        return {k.replace("module.encoder_q.", ""): v for k, v in weights["state_dict"].items()
                if k.startswith("module.encoder_q.") and not k.startswith("module.encoder_q.fc")}
    if architecture in ("swav", "deepcluster_v2"):
        return {k.replace("module.", ""): v for k, v in weights.items()
                if k.startswith("module.") and not k.startswith("module.pro")}
    return weights                                            # dino


class DilationFeatureExtractor(nn.Module):
    """models/image_model.py:90-143.  `image_weights`: None or a state_dict for the encoder (the reference downloads
    dino / moco / swav weights with `requests`, image_model.py:38-44 -- out of scope offline; pass the adapted dict)."""

    def __init__(self, image_weights=None, preprocessing=None):
        super().__init__()
        self.encoder = ResNetEncoder(block=Bottleneck, layers=[3, 4, 6, 3],
                                     replace_stride_with_dilation=[True, True, True])
        if isinstance(image_weights, dict):
            self.encoder.load_state_dict(dict(image_weights))
        elif image_weights == 'imagenet':
            raise FileNotFoundError("image_weights='imagenet' downloads torchvision's checkpoint in the reference "
                                    "(image_model.py:107-108): load it yourself and pass the state_dict")
        else:
            weights = adapt_weights(image_weights)             # :110-113
            if weights is not None:
                self.encoder.load_state_dict(weights)
                print("Loaded '{}' weights for the teacher network~".format(image_weights))
        for param in self.encoder.parameters():
            param.requires_grad = False
        self.decoder = nn.Sequential(nn.Conv2d(2048, 256, 1),
                                     nn.Upsample(scale_factor=4, mode="bilinear", align_corners=True))
        self.preprocessing = preprocessing
        self.normalize_feature = True
        self.channel_avgpooling = nn.AvgPool2d((32, 1), stride=(32, 1))
        self.upsample4 = nn.Upsample(scale_factor=4, mode="bilinear", align_corners=True)

    def forward(self, x):
        if self.preprocessing:
            x = self.preprocessing(x)
        x = self.encoder(x)                                   # [B, 2048, H/4, W/4]
        x = self.decoder(x)                                   # [B, 256, H, W]
        if self.normalize_feature:
            x = F.normalize(x, p=2, dim=1)
        return x

    def forward_pooled(self, x, superpixels, superpixel_size, M):
        """Fused training path: q [M, 256] = superpixel mean-pool of forward(x) (pretrain_trainer.py:434 + :461-463)
        without materialising the [B, 256, H, W] map: encoder -> decoder 1x1 conv (autograd) -> `oess_upnorm_pool`."""
        if self.preprocessing:
            x = self.preprocessing(x)
        feats = self.encoder(x)                               # [B, 2048, H/4, W/4]
        conv = self.decoder[0]
        if USE_TENSOR_CORES and feats.is_cuda and feats.dtype == torch.float32:
            # trainable 1x1 conv 2048 -> 256: tcgen05 forward + backward-weight (the frozen encoder needs no backward-data)
            d = _tc.conv2d_tc_autograd(feats, conv.weight, conv.bias, 0, 1)
        else:
            d = conv(feats)                                   # [B, 256, H/4, W/4]
        return _tc.upnorm_pool(d, superpixels, superpixel_size, M, scale=4)
