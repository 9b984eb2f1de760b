"""E2VID recurrent event encoder, latent-only and sync-free.

Mirror of the reference's e2vid/model/{model.py:69-100 (E2VIDRecurrent), unet.py:108-170 (UNetRecurrent),
submodules.py (ConvLayer, RecurrentConvLayer, ConvLSTM, ResidualBlock, TransposedConvLayer)} with the SAME module
tree and parameter names, so a reference checkpoint (`E2VID_lightweight.pth.tar` -> `state_dict`) loads with
`load_state_dict(strict=True)`.

What changes on the B200:
  * every OpenESS trainer consumes only `latent` (unet.py:163, pretrain_trainer.py:441): with `latent_only=True`
    (default) the three up-sampling decoders, `pred` and the sigmoid are skipped (they stay in the module tree for
    checkpoint compatibility) -- the image is returned as None;
  * the ConvLSTM pointwise tail (3 sigmoids, 2 tanh, cell/hidden update: ~10 ATen kernels per level and step) is one
    fused kernel, `oess_convlstm_gates`;
  * eval-mode BatchNorm of the frozen encoder is folded into the preceding conv once (`fold_bn()`);
  * the ConvLSTM step (cat + 3x3 Gates conv + pointwise tail = 65 % of the encoder's FLOPs) is ONE tensor-core
    kernel, `oess_convlstm_step_nhwc` (tcgen05 implicit GEMM, TF32 operands / fp32 accumulate, channels-last), for the
    hidden sizes of the real E2VID config (multiples of 64); `USE_TENSOR_CORES = False` (or OESS_E2VID_TC=0) keeps the
    strict-fp32 cuDNN + fused-gates path;
  * the strided 5x5 encoder convolutions (folded BN + ReLU fused in the epilogue) run on the same tensor-core
    skeleton, `oess_conv2d_nhwc_tf32` (strided 4-D TMA boxes); the 5-channel head conv (Cin = 5: too thin for a
    128-byte TMA row) goes through a repack to channels-last with zero-padded channels (`oess_planes_to_nhwc_padded`) and
    the same kernel: the whole latent-only encoder runs on hand-written kernels.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import losses as _ops
from ... import ops as _tc

USE_TENSOR_CORES = os.environ.get("OESS_E2VID_TC", "1") != "0"
# bf16 operands (tcgen05.mma.kind::f16, fp32 accumulate / cell state) for the ConvLSTM steps of the FROZEN encoder: the strided
# encoder convolution writes its output as bf16, the ConvLSTM keeps a bf16 copy of its hidden state for the next step.
# OESS_E2VID_DTYPE=tf32 keeps fp32 operands read as TF32.
CONVLSTM_BF16 = os.environ.get("OESS_E2VID_DTYPE", "bf16") == "bf16"


class ConvLayer(nn.Module):
    """submodules.py:7-31."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, activation='relu', norm=None):
        super().__init__()
        bias = False if norm == 'BN' else True
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self.activation = getattr(torch, activation, 'relu') if activation is not None else None
        self.norm = norm
        if norm == 'BN':
            self.norm_layer = nn.BatchNorm2d(out_channels)
        elif norm == 'IN':
            self.norm_layer = nn.InstanceNorm2d(out_channels, track_running_stats=True)
        self._folded = None
        self._packed = None

    def _tc_ok(self, x):
        c = self.conv2d
        return (USE_TENSOR_CORES and x.is_cuda and not torch.is_grad_enabled() and not self.training
                and c.in_channels % 4 == 0 and c.in_channels >= 16 and c.groups == 1 and c.padding_mode == 'zeros'
                and c.kernel_size[0] == c.kernel_size[1] and c.stride[0] == c.stride[1] and c.padding[0] == c.padding[1]
                and c.dilation[0] == c.dilation[1] and (self.norm is None or self._folded is not None)
                and self.activation in (None, torch.relu))

    def _tc_weights(self):
        w, b = self._folded if self._folded is not None else (self.conv2d.weight, self.conv2d.bias)
        key = (w.data_ptr(), w._version, w.device)
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, _tc.conv2d_pack(w), None if b is None else b.detach().float().contiguous())
        return self._packed[1], self._packed[2]

    def _tc_weights_bf16(self):
        """bf16 packing for `oess_conv2d_nhwc_bf16` (input available as bf16: the hidden state of a bf16 ConvLSTM step)."""
        w, b = self._folded if self._folded is not None else (self.conv2d.weight, self.conv2d.bias)
        key = (w.data_ptr(), w._version, w.device)
        if getattr(self, "_packed_bf16", None) is None or self._packed_bf16[0] != key:
            self._packed_bf16 = (key, _tc.conv2d_pack_bf16(w), None if b is None else b.detach().float().contiguous())
        return self._packed_bf16[1], self._packed_bf16[2]

    def fold_bn(self):
        """Eval-mode BN folded into the conv: w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps).
        A later `load_state_dict`, `.to(device)` / `.cuda()` / `.float()` or `train()` marks the fold stale (`_apply`,
        `_load_from_state_dict`, `train` below) and it is redone lazily at the next forward, so the order of fold / load / move does
        not matter; no per-forward key comparison: the E2VID step is launch-bound on the host side."""
        if self.norm == 'BN' and not self.training:
            bn = self.norm_layer
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            self._folded = ((self.conv2d.weight * scale[:, None, None, None]).detach(),
                            (bn.bias - bn.running_mean * scale).detach())
            self._fold_stale = False

    def _refresh_fold(self):
        if self._folded is not None and getattr(self, "_fold_stale", False):
            if self.training:
                self._folded = None
            else:
                self.fold_bn()

    def _apply(self, fn, *args, **kwargs):
        self._fold_stale = True
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._fold_stale = True
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        if mode != self.training:
            self._fold_stale = True
        return super().train(mode)

    def forward(self, x):
        self._refresh_fold()
        if self._tc_ok(x):
            wp, b = self._tc_weights()
            c = self.conv2d
            return _tc.conv2d_tc(x, wp, b, c.kernel_size[0], c.stride[0], c.padding[0], c.dilation[0],
                                 relu=self.activation is not None)
        if self._folded is not None and not self.training:
            out = F.conv2d(x, self._folded[0], self._folded[1], self.conv2d.stride, self.conv2d.padding)
        else:
            out = self.conv2d(x)
            if self.norm in ['BN', 'IN']:
                out = self.norm_layer(out)
        if self.activation is not None:
            out = self.activation(out)
        return out


class TransposedConvLayer(nn.Module):
    """submodules.py:34-63."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, activation='relu', norm=None):
        super().__init__()
        bias = False if norm == 'BN' else True
        self.transposed_conv2d = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=2, padding=padding,
                                                    output_padding=1, bias=bias)
        self.activation = getattr(torch, activation, 'relu') if activation is not None else None
        self.norm = norm
        if norm == 'BN':
            self.norm_layer = nn.BatchNorm2d(out_channels)
        elif norm == 'IN':
            self.norm_layer = nn.InstanceNorm2d(out_channels, track_running_stats=True)

    def forward(self, x):
        out = self.transposed_conv2d(x)
        if self.norm in ['BN', 'IN']:
            out = self.norm_layer(out)
        if self.activation is not None:
            out = self.activation(out)
        return out


class ConvLSTM(nn.Module):
    """submodules.py:175-214; the pointwise tail is the fused CUDA kernel when running without grad on the GPU."""

    def __init__(self, input_size, hidden_size, kernel_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.Gates = nn.Conv2d(input_size + hidden_size, 4 * hidden_size, kernel_size, padding=kernel_size // 2)
        self._packed = None

    def _tc_weights(self):
        """Gates.weight / bias repacked for the tensor-core kernel; rebuilt when the parameters change or move."""
        w, b = self.Gates.weight, self.Gates.bias
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, w.device)
        if self._packed is None or self._packed[0] != key:
            self._packed = (key,) + _tc.convlstm_pack(w, b, self.hidden_size)
        return self._packed[1], self._packed[2]

    def forward(self, input_, prev_state=None):
        if (USE_TENSOR_CORES and input_.is_cuda and not torch.is_grad_enabled() and self.hidden_size % 64 == 0
                and self.input_size == self.hidden_size and tuple(self.Gates.kernel_size) == (3, 3)):
            wp, bp = self._tc_weights()
            if CONVLSTM_BF16:
                if getattr(self, "_packed_bf", None) is None or self._packed_bf[0] is not wp:
                    self._packed_bf = (wp, wp.to(torch.bfloat16))
                return _tc.convlstm_step_bf16(input_, prev_state, self._packed_bf[1], bp)
            return _tc.convlstm_step(input_, prev_state, wp, bp)
        if prev_state is None:
            # zero state: the hidden half of the stacked input contributes nothing -> convolve the input half only
            w = self.Gates.weight[:, :self.input_size]
            gates = F.conv2d(input_, w, self.Gates.bias, padding=self.Gates.padding)
            prev_cell = None
        else:
            prev_hidden, prev_cell = prev_state
            gates = self.Gates(torch.cat((input_, prev_hidden), 1))
        if gates.is_cuda and not torch.is_grad_enabled():
            return _ops.convlstm_gates(gates, prev_cell)
        in_gate, remember_gate, out_gate, cell_gate = gates.chunk(4, 1)      # differentiable path (unfrozen E2VID)
        prev_c = 0 if prev_cell is None else prev_cell
        cell = torch.sigmoid(remember_gate) * prev_c + torch.sigmoid(in_gate) * torch.tanh(cell_gate)
        hidden = torch.sigmoid(out_gate) * torch.tanh(cell)
        return hidden, cell


class RecurrentConvLayer(nn.Module):
    """submodules.py:96-115 (convlstm only: the only recurrent block type OpenESS uses)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, recurrent_block_type='convlstm',
                 activation='relu', norm=None):
        super().__init__()
        assert recurrent_block_type == 'convlstm', "OpenESS runs E2VID with ConvLSTM blocks"
        self.recurrent_block_type = recurrent_block_type
        self.conv = ConvLayer(in_channels, out_channels, kernel_size, stride, padding, activation, norm)
        self.recurrent_block = ConvLSTM(input_size=out_channels, hidden_size=out_channels, kernel_size=3)

    def forward(self, x, prev_state):
        rb = self.recurrent_block
        self.conv._refresh_fold()
        if (CONVLSTM_BF16 and USE_TENSOR_CORES and self.conv._tc_ok(x) and rb.hidden_size % 64 == 0
                and rb.input_size == rb.hidden_size and tuple(rb.Gates.kernel_size) == (3, 3)):
            c = self.conv.conv2d                                # the conv output goes straight to bf16: ConvLSTM operand
            xb = getattr(x, "_oess_bf16", None)                 # bf16 copy of the previous level's hidden state
            if xb is not None and c.in_channels % 64 == 0:
                wp, b = self.conv._tc_weights_bf16()
                x = _tc.conv2d_tc_bf16(xb, wp, b, c.kernel_size[0], c.stride[0], c.padding[0], c.dilation[0],
                                       relu=self.conv.activation is not None, want_f32=False)[1]
            else:
                wp, b = self.conv._tc_weights()
                x = _tc.conv2d_tc_bf16out(x, wp, b, c.kernel_size[0], c.stride[0], c.padding[0], c.dilation[0],
                                          relu=self.conv.activation is not None)
        else:
            x = self.conv(x)
        state = self.recurrent_block(x, prev_state)
        return state[0], state


class ResidualBlock(nn.Module):
    """submodules.py:140-172."""

    def __init__(self, in_channels, out_channels, stride=1, downsample=None, norm=None):
        super().__init__()
        bias = False if norm == 'BN' else True
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=bias)
        self.norm = norm
        if norm == 'BN':
            self.bn1 = nn.BatchNorm2d(out_channels)
            self.bn2 = nn.BatchNorm2d(out_channels)
        elif norm == 'IN':
            self.bn1 = nn.InstanceNorm2d(out_channels)
            self.bn2 = nn.InstanceNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=bias)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.conv1(x)
        if self.norm in ['BN', 'IN']:
            out = self.bn1(out)
        out = self.relu(out)
        out = self.conv2(out)
        if self.norm in ['BN', 'IN']:
            out = self.bn2(out)
        if self.downsample:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class UNetRecurrent(nn.Module):
    """unet.py:16-63 (BaseUNet) + :108-170 (UNetRecurrent), skip_type 'sum' or 'concat', TransposedConv decoders."""

    def __init__(self, num_input_channels, num_output_channels=1, skip_type='sum', recurrent_block_type='convlstm',
                 activation='sigmoid', num_encoders=4, base_num_channels=32, num_residual_blocks=2, norm=None,
                 use_upsample_conv=True):
        super().__init__()
        if use_upsample_conv:
            raise NotImplementedError("UpsampleConvLayer decoders are dead code on the OpenESS latent path; the "
                                      "lightweight checkpoint uses TransposedConvLayer")
        self.skip_type = skip_type
        self.num_encoders = num_encoders
        self.norm = norm
        self.activation = getattr(torch, activation, 'sigmoid')
        c = base_num_channels
        self.max_num_channels = c * pow(2, num_encoders)
        self.head = ConvLayer(num_input_channels, c, kernel_size=5, stride=1, padding=2)
        self.encoders = nn.ModuleList()
        for i in range(num_encoders):
            self.encoders.append(RecurrentConvLayer(c * pow(2, i), c * pow(2, i + 1), kernel_size=5, stride=2, padding=2,
                                                    recurrent_block_type=recurrent_block_type, norm=norm))
        self.resblocks = nn.ModuleList([ResidualBlock(self.max_num_channels, self.max_num_channels, norm=norm)
                                        for _ in range(num_residual_blocks)])
        self.decoders = nn.ModuleList()
        for input_size in reversed([c * pow(2, i + 1) for i in range(num_encoders)]):
            self.decoders.append(TransposedConvLayer(input_size if skip_type == 'sum' else 2 * input_size,
                                                     input_size // 2, kernel_size=5, padding=2, norm=norm))
        self.pred = ConvLayer(c if skip_type == 'sum' else 2 * c, num_output_channels, 1, activation=None, norm=norm)

    def _skip(self, a, b):
        return a + b if self.skip_type == 'sum' else torch.cat([a, b], dim=1)

    def _head_tc(self, x):
        """Head conv (Cin = num_bins = 5: too thin for a 16-byte TMA row) on the tensor cores: the input planes are
        repacked to channels-last with the channels zero-padded to 8, the weights get matching zero channels."""
        conv = self.head.conv2d
        w, b = conv.weight, conv.bias
        key = (w.data_ptr(), w._version, w.device)
        k = conv.kernel_size[0]
        # odd 'same' kernels: the taps of a kernel row are folded into the channel dimension (oess_conv2d_nhwc_tf32_rowunfold:
        # K = 5 x 64 instead of 25 x 32 for the 5 x 5 head of E2VID)
        unfold = (conv.kernel_size[0] == conv.kernel_size[1] and k % 2 == 1 and conv.stride[0] == 1 and conv.dilation[0] == 1
                  and conv.padding[0] == k // 2 and k * 8 <= 256)
        if getattr(self, "_head_packed", None) is None or self._head_packed[0] != key:
            wpad = torch.zeros(w.shape[0], 8, w.shape[2], w.shape[3], dtype=torch.float32, device=w.device)
            wpad[:, :w.shape[1]] = w.detach()
            self._head_packed = (key, _tc.conv2d_pack_rowunfold(wpad) if unfold else _tc.conv2d_pack(wpad),
                                 None if b is None else b.detach().float().contiguous())
        if unfold:
            x8 = _tc.planes_to_nhwc_padded_w(x, 8, k // 2)
            return _tc.conv2d_rowunfold(x8, self._head_packed[1], self._head_packed[2], k, k, x.shape[3],
                                        relu=self.head.activation is not None)
        x8 = _tc.planes_to_nhwc_padded(x, 8)
        return _tc.conv2d_tc(x8, self._head_packed[1], self._head_packed[2], conv.kernel_size[0], conv.stride[0],
                             conv.padding[0], conv.dilation[0], relu=self.head.activation is not None)

    # ---- image branch on own kernels (online reconstruction instead of precomputed PNGs) ----
    def _decode_tc_ok(self, x):
        chans = [d.transposed_conv2d.in_channels for d in self.decoders] + [d.transposed_conv2d.out_channels for d in self.decoders]
        return (USE_TENSOR_CORES and x.is_cuda and not torch.is_grad_enabled() and not self.training and self.skip_type == 'sum'
                and self.norm in (None, 'BN') and all(c % 4 == 0 for c in chans) and self.pred.conv2d.in_channels <= 64
                and self.pred.conv2d.out_channels == 1 and self.activation is torch.sigmoid
                and all(d.activation in (None, torch.relu) for d in self.decoders))

    @staticmethod
    def _bn_fold(bn, bias, n, device):
        """(scale, shift) of an eval-mode BatchNorm (identity if bn is None), the conv bias folded in."""
        if bn is None:
            scale, shift = torch.ones(n, device=device), torch.zeros(n, device=device)
        else:
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float()
            shift = (bn.bias - bn.running_mean * scale).detach().float()
        if bias is not None:
            shift = shift + bias.detach().float() * scale
        return scale, shift

    def _decode_tc(self, x, blocks, head):
        """resblocks -> decoders (skip sum fused into the zero insertion; ConvTranspose2d = tcgen05 conv over the zero-inserted
        map with the rotated kernel, BN folded, ReLU in the epilogue) -> prediction layer + sigmoid (unet.py:165-168)."""
        from ...models import _tc_resnet as _tcr
        cache = self.__dict__.setdefault("_dec_cache", _tcr.PackedConvCache())
        packed = self.__dict__.setdefault("_dec_packed", {})
        for rb in self.resblocks:                          # submodules.py:140-172
            bn1, bn2 = (rb.bn1, rb.bn2) if self.norm == 'BN' else (None, None)
            out = _tcr.conv_bn(cache, x, rb.conv1, bn1, True)
            x = _tcr.conv_bn(cache, out, rb.conv2, bn2, True, residual=x if rb.downsample is None else rb.downsample(x))
        for i, dec in enumerate(self.decoders):            # submodules.py:34-63
            tc = dec.transposed_conv2d
            w = tc.weight
            key = (w.data_ptr(), w._version, w.device)
            hit = packed.get(i)
            if hit is None or hit[0] != key:
                scale, shift = self._bn_fold(dec.norm_layer if dec.norm == 'BN' else None, tc.bias, tc.out_channels, w.device)
                hit = (key, _tc.conv_transpose2x_pack(w, scale), shift.contiguous())
                packed[i] = hit
            z = _tc.zero_insert2x_nhwc(x, blocks[self.num_encoders - i - 1])
            k, p = tc.kernel_size[0], tc.padding[0]
            x = _tc.conv2d_tc(z, hit[1], hit[2], k, 1, k - 1 - p, 1, relu=dec.activation is not None, round_out=True)
        pc = self.pred.conv2d
        key = (pc.weight.data_ptr(), pc.weight._version, pc.weight.device)
        hit = packed.get("pred")
        if hit is None or hit[0] != key:
            scale, shift = self._bn_fold(self.pred.norm_layer if self.pred.norm == 'BN' else None, pc.bias, 1, pc.weight.device)
            hit = (key, (pc.weight.detach().float().reshape(-1) * scale).contiguous(), float(shift))
            packed["pred"] = hit
        return _tc.pred_sigmoid_nhwc(x, head, hit[1], hit[2])

    def forward(self, x, prev_states, latent_only=True):
        hc = self.head.conv2d
        if (USE_TENSOR_CORES and x.is_cuda and not torch.is_grad_enabled() and self.head.norm is None
                and hc.in_channels <= 8 and hc.out_channels % 4 == 0 and hc.out_channels >= 16
                and self.head.activation in (None, torch.relu)):
            x = self._head_tc(x)
        else:
            x = self.head(x)
        head = x
        if prev_states is None:
            prev_states = [None] * self.num_encoders
        blocks, states = [], []
        for i, encoder in enumerate(self.encoders):
            x, state = encoder(x, prev_states[i])
            blocks.append(x)
            states.append(state)
        latent = {1: head}
        for i in range(self.num_encoders):
            latent[2 ** (i + 1)] = blocks[i]
        if self.num_encoders > 3:                       # unet.py:163 keeps exactly {1, 2, 4, 8}
            latent = {k: latent[k] for k in (1, 2, 4, 8)}
        if latent_only:
            return None, states, latent                 # resblocks + decoders + pred feed only the image
        if self._decode_tc_ok(x):
            return self._decode_tc(x, blocks, head), states, latent
        for resblock in self.resblocks:
            x = resblock(x)
        for i, decoder in enumerate(self.decoders):
            x = decoder(self._skip(x, blocks[self.num_encoders - i - 1]))
        img = self.activation(self.pred(self._skip(x, head)))
        return img, states, latent


class E2VIDRecurrent(nn.Module):
    """model.py:69-100 with the config keys of model.py:13-46 (same defaults)."""

    def __init__(self, config, latent_only=True):
        super().__init__()
        assert 'num_bins' in config
        self.num_bins = int(config['num_bins'])
        self.skip_type = str(config.get('skip_type', 'sum'))
        self.num_encoders = int(config.get('num_encoders', 4))
        self.base_num_channels = int(config.get('base_num_channels', 32))
        self.num_residual_blocks = int(config.get('num_residual_blocks', 2))
        self.norm = str(config['norm']) if 'norm' in config else None
        self.use_upsample_conv = bool(config.get('use_upsample_conv', True))
        self.recurrent_block_type = str(config.get('recurrent_block_type', 'convlstm'))
        self.latent_only = latent_only
        self.unetrecurrent = UNetRecurrent(num_input_channels=self.num_bins, num_output_channels=1,
                                           skip_type=self.skip_type, recurrent_block_type=self.recurrent_block_type,
                                           activation='sigmoid', num_encoders=self.num_encoders,
                                           base_num_channels=self.base_num_channels,
                                           num_residual_blocks=self.num_residual_blocks, norm=self.norm,
                                           use_upsample_conv=self.use_upsample_conv)

    def fold_bn(self):
        for m in self.modules():
            if isinstance(m, ConvLayer):
                m.fold_bn()
        return self

    def forward(self, event_tensor, prev_states):
        return self.unetrecurrent.forward(event_tensor, prev_states, latent_only=self.latent_only)
