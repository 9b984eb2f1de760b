"""Mirror of the reference's models/style_networks.py:SemSegE2VID (task decoder on the E2VID latents)
for the configuration every OpenESS trainer uses (skip_connect=True, skip_type='concat'): SAME module tree and
state_dict keys (decoder_scale_1..5, decoder_ch256, decoder_ch512, text_embeddings, linear_probe), so reference
checkpoints load with strict=True.

What changes on the B200: the head `decoder_ch256 -> decoder_ch512 -> conv(text_embeddings)` (style_networks.py:163-165)
is affine with no non-linearity.  The logits are computed directly from the 32-channel map with the collapsed weight
W_eff = T W512 W256 (one HBM-bound `oess_pixel_linear` pass, autograd-exact chain rule through the tiny matrices), so the
512-channel full-resolution map (4.6 GB at batch 8) never exists.  `forward()` keeps the reference signature
(it still returns the 256-channel map x_ch256); `forward_pooled()` is the fully fused training path: it returns the
superpixel-pooled contrastive features k directly (pool the 32-channel map, then apply decoder_ch256 to the [M, 32]
means -- exact by linearity), so the 2.3 GB x_ch256 map never exists either.
Forward-only uses (validation / test.py under no_grad, linear probing with the trunk frozen) run the trunk on the tensor
cores: conv + InstanceNorm (+ residual) (+ ReLU) = `oess_conv2d_nhwc_tf32_instats` (per-sample statistics in the TMEM
epilogue) + `oess_instancenorm_nhwc_sums`.  Under autograd (pretraining: the trunk trains) every block is `ops.conv_in_autograd`: tcgen05 forward conv with
per-sample statistics, InstanceNorm Jacobian kernels, tcgen05 backward-data (forward kernel on rotated weights) and
backward-weight (split-K MN-major GEMM) convolutions -- no cuDNN call is left in the module
(`TRAIN_ON_TENSOR_CORES = False` / OESS_SEMSEG_TC_TRAIN=0 restores the strict-fp32 torch path)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as f

from .. import ops as _ops
from ..losses import segpool_forward, superpixel_pool  # noqa: F401

USE_TENSOR_CORES = os.environ.get("OESS_SEMSEG_TC", "1") != "0"
TRAIN_ON_TENSOR_CORES = os.environ.get("OESS_SEMSEG_TC_TRAIN", "1") != "0"   # training trunk (fwd + bwd) on own kernels


def skip_concat(x1, x2):
    return torch.cat([x1, x2], dim=1)


def skip_sum(x1, x2):
    return x1 + x2


def gaussian_weights_init(m):
    classname = m.__class__.__name__
    if classname.find('Conv') != -1 and classname.find('Conv') == 0:
        m.weight.data.normal_(0.0, 0.02)


class ReLUINSConv2d(nn.Module):
    """style_networks.py:252-263."""

    def __init__(self, n_in, n_out, kernel_size, stride, padding=0):
        super().__init__()
        self.model = nn.Sequential(nn.Conv2d(n_in, n_out, kernel_size=kernel_size, stride=stride, padding=padding, bias=True),
                                   nn.InstanceNorm2d(n_out, affine=False), nn.ReLU(inplace=True))
        self.model.apply(gaussian_weights_init)

    def forward(self, x):
        return self.model(x)


class INSResBlock(nn.Module):
    """style_networks.py:266-289."""

    def __init__(self, inplanes, planes, stride=1, dropout=0.0):
        super().__init__()
        model = [nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1), nn.InstanceNorm2d(planes),
                 nn.ReLU(inplace=True), nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1),
                 nn.InstanceNorm2d(planes)]
        if dropout > 0:
            model += [nn.Dropout(p=dropout)]
        self.model = nn.Sequential(*model)
        self.model.apply(gaussian_weights_init)

    def forward(self, x):
        return self.model(x) + x


class SemSegE2VID(nn.Module):
    def __init__(self, input_c, output_c, skip_connect=False, skip_type='sum', input_index_map=False,
                 text_embeddings_path='', if_linear_probing=False):
        super().__init__()
        if not skip_connect or input_index_map:
            raise NotImplementedError("OpenESS builds SemSegE2VID with skip_connect=True, input_index_map=False "
                                      "(training/pretrain_trainer.py:174-180)")
        self.skip_connect = skip_connect
        self.skip_type = skip_type
        self.apply_skip_connection = skip_sum if skip_type == 'sum' else skip_concat
        tch = input_c
        self.text_embeddings_path = text_embeddings_path
        if text_embeddings_path is None:
            self.text_embeddings = nn.Parameter(torch.zeros(output_c, 512))
            nn.init.normal_(self.text_embeddings, mean=0.0, std=0.01)
        else:
            self.register_buffer('text_embeddings', torch.randn(output_c, 512))
            if text_embeddings_path:
                loaded = torch.load(text_embeddings_path, map_location='cpu')
                self.text_embeddings[:, :] = loaded[:, :]
        self.decoder_scale_1 = nn.Sequential(*([INSResBlock(tch, tch) for _ in range(5)] +
                                               [ReLUINSConv2d(tch, tch // 2, kernel_size=3, stride=1, padding=1)]))
        self.decoder_scale_2 = nn.Sequential(ReLUINSConv2d(tch, tch // 2, kernel_size=3, stride=1, padding=1),
                                             ReLUINSConv2d(tch // 2, tch // 4, kernel_size=3, stride=1, padding=1))
        tch = tch // 2
        self.decoder_scale_3 = nn.Sequential(ReLUINSConv2d(tch, tch // 2, kernel_size=3, stride=1, padding=1),
                                             ReLUINSConv2d(tch // 2, tch // 2, kernel_size=3, stride=1, padding=1))
        tch = tch // 2
        self.decoder_scale_4 = nn.Sequential(ReLUINSConv2d(tch, tch // 2, kernel_size=3, stride=1, padding=1))
        tch = tch // 2
        self.head_in = tch
        self.decoder_scale_5 = nn.Sequential(nn.Conv2d(tch, output_c, kernel_size=1, stride=1, padding=0))   # unused (:167)
        self.decoder_ch256 = nn.Sequential(nn.Conv2d(tch, 256, kernel_size=1, stride=1, padding=0))
        self.decoder_ch512 = nn.Sequential(nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0))
        self._pack_cache = {}
        self.if_linear_probing = if_linear_probing
        if if_linear_probing:
            for blk in (self.decoder_scale_1, self.decoder_scale_2, self.decoder_scale_3, self.decoder_scale_4,
                        self.decoder_ch256, self.decoder_ch512):
                for p in blk.parameters():
                    p.requires_grad = False
            self.linear_probe = nn.Conv2d(output_c, output_c, 1)

    def update_skip_dict(self, skips, x, sz_in):
        rem, scale = sz_in % x.shape[3], sz_in // x.shape[3]
        assert rem == 0
        skips[scale] = x

    # ---- trunk: decoder_scale_1..4 (style_networks.py:146-160) -> 32-channel full-resolution map
    def trunk(self, input_dict, out):
        if self._tc_trunk_ok(input_dict):
            return self.trunk_tc(input_dict, out)
        if self._tc_train_ok(input_dict):
            return self.trunk_tc(input_dict, out, train=True)
        sz_in = input_dict[1].shape[3]
        x = self.decoder_scale_1(input_dict[8])
        x = f.interpolate(x, scale_factor=2, mode='nearest')
        x = self.apply_skip_connection(x, input_dict[4])
        x = self.decoder_scale_2(x)
        self.update_skip_dict(out, x, sz_in)
        x = f.interpolate(x, scale_factor=2, mode='nearest')
        x = self.apply_skip_connection(x, input_dict[2])
        x = self.decoder_scale_3(x)
        self.update_skip_dict(out, x, sz_in)
        x = f.interpolate(x, scale_factor=2, mode='nearest')
        return self.decoder_scale_4(x)

    # ---- trunk on the tensor cores, forward only (validation, test.py, linear probing: no gradient reaches the trunk)
    def _tc_trunk_ok(self, input_dict):
        x = input_dict[8]
        trunk_frozen = not any(p.requires_grad for blk in (self.decoder_scale_1, self.decoder_scale_2, self.decoder_scale_3,
                                                            self.decoder_scale_4) for p in blk.parameters())
        return (USE_TENSOR_CORES and x.is_cuda and x.dtype == torch.float32 and self.skip_type == 'concat'
                and (not torch.is_grad_enabled() or trunk_frozen) and not any(v.requires_grad for v in input_dict.values())
                and all(v.shape[1] % 4 == 0 for v in input_dict.values()))

    def _packed(self, conv):
        key = (conv.weight.data_ptr(), conv.weight._version, conv.weight.device)
        hit = self._pack_cache.get(id(conv))
        if hit is None or hit[0] != key:
            hit = (key, _ops.conv2d_pack(conv.weight), None if conv.bias is None else conv.bias.detach().float().contiguous())
            self._pack_cache[id(conv)] = hit
        return hit[1], hit[2]

    def _conv_in(self, x, conv, norm, relu, residual=None, train=False):
        if train:                                                 # differentiable: fwd + dgrad + wgrad + IN Jacobian
            return _ops.conv_in_autograd(x, conv.weight, conv.bias, residual, padding=conv.padding[0],
                                         dilation=conv.dilation[0], eps=norm.eps, relu=relu)
        wp, b = self._packed(conv)
        return _ops.conv_in(x, wp, b, conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0], eps=norm.eps,
                            residual=residual, relu=relu)

    def _seq_tc(self, seq, x, train=False):
        for blk in seq:
            if isinstance(blk, INSResBlock):                      # conv-IN-ReLU-conv-IN, + x (style_networks.py:266-289)
                m = blk.model
                h = self._conv_in(x, m[0], m[1], True, train=train)
                x = self._conv_in(h, m[3], m[4], False, residual=x, train=train)
            else:                                                 # ReLUINSConv2d: conv-IN-ReLU (:252-263)
                m = blk.model
                x = self._conv_in(x, m[0], m[1], True, train=train)
        return x

    def trunk_tc(self, input_dict, out, train=False):
        """`trunk` with every conv + InstanceNorm (+ residual) (+ ReLU) on the hand-written kernels; nearest x2 upsampling
        + skip concatenation is one kernel (`ops.upsample2x_cat`).  train=False: forward only (no autograd graph);
        train=True: differentiable blocks (`ops.conv_in_autograd`)."""
        with torch.set_grad_enabled(train and torch.is_grad_enabled()):
            cl = torch.channels_last
            sz_in = input_dict[1].shape[3]
            x = self._seq_tc(self.decoder_scale_1, input_dict[8], train)
            x = _ops.upsample2x_cat(x, input_dict[4])            # nearest x2 + skip concat in one pass (:148-150)
            x = self._seq_tc(self.decoder_scale_2, x, train)
            self.update_skip_dict(out, x, sz_in)
            x = _ops.upsample2x_cat(x, input_dict[2])
            x = self._seq_tc(self.decoder_scale_3, x, train)
            self.update_skip_dict(out, x, sz_in)
            x = _ops.upsample2x_cat(x)
            return self._seq_tc(self.decoder_scale_4, x, train)

    def _tc_train_ok(self, input_dict):
        x = input_dict[8]
        convs = [m for blk in (self.decoder_scale_1, self.decoder_scale_2, self.decoder_scale_3, self.decoder_scale_4)
                 for m in blk.modules() if isinstance(m, nn.Conv2d)]
        return (USE_TENSOR_CORES and TRAIN_ON_TENSOR_CORES and x.is_cuda and x.dtype == torch.float32
                and self.skip_type == 'concat' and torch.is_grad_enabled()
                and all(v.shape[1] % 4 == 0 for v in input_dict.values())
                and all(c.stride == (1, 1) and c.kernel_size[0] == c.kernel_size[1] and c.out_channels % 4 == 0
                        and (256 % (c.out_channels // 4) == 0 or (c.out_channels // 4) % 256 == 0) for c in convs))

    # ---- collapsed head weights (tiny, differentiable w.r.t. decoder_ch256 / decoder_ch512 / text_embeddings / linear_probe)
    def collapsed_head(self):
        W256 = self.decoder_ch256[0].weight.flatten(1)            # [256, 32]
        b256 = self.decoder_ch256[0].bias
        W512 = self.decoder_ch512[0].weight.flatten(1)            # [512, 256]
        b512 = self.decoder_ch512[0].bias
        T = self.text_embeddings                                   # [K, 512]
        TW = T @ W512                                              # [K, 256]
        W_eff = TW @ W256                                          # [K, 32]
        b_eff = TW @ b256 + T @ b512                               # [K]
        if self.if_linear_probing:
            L = self.linear_probe.weight.flatten(1)                # [K, K]
            W_eff, b_eff = L @ W_eff, L @ b_eff + self.linear_probe.bias
        return W_eff, b_eff

    def logits_from_trunk(self, x32):
        W_eff, b_eff = self.collapsed_head()
        if x32.is_cuda and self.head_in <= 64 and W_eff.shape[0] <= 64:
            return _ops.pixel_linear(x32, W_eff, b_eff)
        return f.conv2d(x32, W_eff[:, :, None, None], b_eff)      # shapes outside the kernel's range

    def forward(self, input_dict):
        """Reference signature: -> (out {8, 4, 2, 1: logits}, x_ch256)."""
        out = {8: input_dict[8]}
        x32 = self.trunk(input_dict, out)
        x_ch256 = self.decoder_ch256(x32)
        self.update_skip_dict(out, self.logits_from_trunk(x32), input_dict[1].shape[3])
        return out, x_ch256

    def forward_pooled(self, input_dict, superpixels, superpixel_size, M=None):
        """Fused training path: -> (out, k) with k = superpixel mean-pool of x_ch256, [M, 256]
        (pretrain_trainer.py:445-459) without materialising x_ch256."""
        out = {8: input_dict[8]}
        x32 = self.trunk(input_dict, out)
        self.update_skip_dict(out, self.logits_from_trunk(x32), input_dict[1].shape[3])
        B = x32.shape[0]
        if M is None:
            off = torch.arange(0, B * superpixel_size, superpixel_size, device=superpixels.device)[:, None, None]
            M = int((superpixels + off).max().item()) + 1
        p32 = superpixel_pool(x32, superpixels, superpixel_size, M)               # [M, 32] = sum / (count + 1e-6)
        with torch.no_grad():
            _, counts = segpool_forward(x32[:, :1].contiguous(), superpixels, superpixel_size, M)
        ratio = counts / (counts + 1e-6)                                          # pooled bias = b * count / (count + 1e-6)
        W256 = self.decoder_ch256[0].weight.flatten(1)
        k = p32 @ W256.t() + ratio[:, None] * self.decoder_ch256[0].bias[None, :]
        return out, k
