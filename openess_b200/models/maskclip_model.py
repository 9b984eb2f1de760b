"""MaskCLIP ViT-B/16 feature extractor: mirror of models/maskclip_model.py
(`maskClipFeatureExtractor` :853-915, `VisionTransformer` :545-851, `TransformerEncoderLayer` :448-541, `PatchEmbed`
:380-445, `MaskClipHead` :52-222) with the SAME module tree and state_dict keys -- `encoder.patch_embed.projection.weight`,
`encoder.cls_token`, `encoder.pos_embed`, `encoder.layers.N.{ln1,ln2}.*`, `encoder.layers.N.attn.attn.{in_proj_weight,
in_proj_bias,out_proj.*}`, `encoder.layers.N.ffn.layers.{0.0,1}.*`, `encoder.{ln0,ln1}.*`, `decoder.{proj.weight,
text_embeddings,image_mapping_local.*}` -- so a MaskCLIP checkpoint converted for the reference
(`load_checkpoint1`, :21-49) loads unchanged.  It does NOT import mmcv / mmseg: the two mmcv bricks the reference uses are
restated by their published definition (mmcv-full 1.6.0, mmcv/cnn/bricks/transformer.py): `MultiheadAttention(batch_first)`
= `nn.MultiheadAttention` + `identity + out`, `FFN` = `Linear -> GELU -> Linear` + identity; dropout / DropPath rates are 0.

The four trainers construct this module and put it in eval() but never call it; it is built module-level.

B200 forward (frozen, no grad): tokens stay row-major [B * T, D] end to end (no [B, L, C] <-> [L, B, C] transposes);
  * patch embedding = `oess_vit_patchify` (zero 'corner' padding folded in) + ONE tcgen05 GEMM;
  * every linear (in_proj, out_proj, fc1 + GELU, fc2, head proj, text classifier) = `oess_gemm_tf32_ex` with bias / GELU /
    residual in the TMEM epilogue -- the residual stream is updated in place by the GEMM that produces the branch;
  * LayerNorm = `oess_layernorm_rows` (one warp per token, the row read once);
  * attention = `oess_mha_fwd_tc` (tcgen05 flash attention: S and the O tile in TMEM, fp32 online softmax, no [T, T] matrix in
    memory; `OESS_MHA=simt` selects the exact-fp32 FMA kernel `oess_mha_fwd`);
  * the last layer's extra value path (`v = out_proj(v_proj(ln1(x))) + x; v = ffn(ln2(v)) + v`, :522-536) reuses the v third
    of the in_proj output the attention needs anyway;
  * head: proj GEMM -> `oess_l2norm_rows` -> classifier GEMM against the text embeddings -> `oess_bilinear_tokens_to_nchw`.
TF32 operands / fp32 accumulate in the GEMMs (the class of torch's default cuDNN / cuBLAS-TF32 convolution path), fp32
everywhere else; stated tolerance on the cosine logits: 2e-3 absolute (measured 1.6e-4 on the full model at 440 x 640, tests/test_maskclip.py).
There is no CPU path: forward raises on a non-CUDA tensor.
"""
import re
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops as _ops
from .._lib import OpenESSB200Error


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def load_checkpoint1(model_load_path, model):
    """models/maskclip_model.py:21-49: shape-matched partial load of `['state_dict']` with the `backbone.` prefix removed."""
    own = model.state_dict()
    pre = torch.load(model_load_path, map_location="cpu")["state_dict"]
    pre = OrderedDict((re.sub(r"^backbone\.", "", k), v) for k, v in pre.items())
    match = nomatch = 0
    for k, v in pre.items():
        if k in own and own[k].shape == v.shape:
            own[k] = v
            match += 1
        else:
            print("missed keys: ", k)
            nomatch += 1
    print("matched parameter sets: {}, and no matched: {}".format(match, nomatch))
    model.load_state_dict(own)
    return model


class _MHAWrapper(nn.Module):
    """mmcv MultiheadAttention: holds `attn = nn.MultiheadAttention` (keys `attn.attn.*`)."""

    def __init__(self, embed_dims, num_heads, bias=True):
        super().__init__()
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout=0.0, bias=bias)


class _FFN(nn.Module):
    """mmcv FFN (num_fcs = 2): `layers = Sequential(Sequential(Linear, GELU, Dropout), Linear, Dropout)`."""

    def __init__(self, embed_dims, feedforward_channels):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.GELU(), nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))


class TransformerEncoderLayer(nn.Module):
    """maskclip_model.py:448-541."""

    def __init__(self, embed_dims, num_heads, feedforward_channels, qkv_bias=True, eps=1e-6):
        super().__init__()
        self.ln1 = nn.LayerNorm(embed_dims, eps=eps)
        self.attn = _MHAWrapper(embed_dims, num_heads, bias=qkv_bias)
        self.ln2 = nn.LayerNorm(embed_dims, eps=eps)
        self.ffn = _FFN(embed_dims, feedforward_channels)

    @property
    def norm1(self):
        return self.ln1

    @property
    def norm2(self):
        return self.ln2

    def _ffn_(self, x):
        """x <- x + ffn(ln2(x)), in place on the [rows, D] residual stream."""
        fc1, fc2 = self.ffn.layers[0][0], self.ffn.layers[1]
        y = _ops.layernorm_rows(x, self.ln2.weight, self.ln2.bias, self.ln2.eps)
        hdn = _ops.gemm_tf32_ex(y, fc1.weight, fc1.bias, act="gelu")
        return _ops.gemm_tf32_ex(hdn, fc2.weight, fc2.bias, residual=x, out=x)

    def forward_rows(self, x, B, T, return_qkv=False):
        """x: [B * T, D] residual stream (updated in place and returned); q, k, v: [B * T, D] or None (:519-541)."""
        mha = self.attn.attn
        D = x.shape[1]
        q = k = v = None
        y = _ops.layernorm_rows(x, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        # q, k, v are tensor-core operands next (attention, out_proj): stored TF32-rounded instead of truncated on read
        qkv = _ops.gemm_tf32_ex(y, mha.in_proj_weight, mha.in_proj_bias, round_out=True)   # [B * T, 3 D]
        if return_qkv:
            # :524-533  y.view(N, L, 3, C).permute(2, 0, 1, 3) -> out_proj on each third; `v += x`; `v = ffn(norm2(v), identity=v)`
            thirds = qkv.view(B * T, 3, D).permute(1, 0, 2).contiguous()
            qk = _ops.gemm_tf32_ex(thirds[:2].reshape(2 * B * T, D), mha.out_proj.weight, mha.out_proj.bias)
            q, k = qk[:B * T], qk[B * T:]
            v = _ops.gemm_tf32_ex(thirds[2], mha.out_proj.weight, mha.out_proj.bias, residual=x)
            v = self._ffn_(v)
        a = _ops.mha_fwd(qkv, B, T, mha.num_heads)
        x = _ops.gemm_tf32_ex(a, mha.out_proj.weight, mha.out_proj.bias, residual=x, out=x)
        x = self._ffn_(x)
        return x, q, k, v

    def forward(self, x, return_qkv=False):
        """Reference signature: x [N, L, C] -> (x, q, k, v)."""
        _require(x)
        N, L, C = x.shape
        with torch.no_grad():
            xr, q, k, v = self.forward_rows(x.reshape(N * L, C).clone(), N, L, return_qkv)
        back = lambda t: None if t is None else t.view(N, L, C)
        return back(xr), back(q), back(k), back(v)


class AdaptivePadding(nn.Module):
    """maskclip_model.py:259-327 ('corner' / 'same' zero padding up to a multiple of the stride)."""

    def __init__(self, kernel_size=1, stride=1, dilation=1, padding="corner"):
        super().__init__()
        assert padding in ("same", "corner")
        self.padding = padding
        self.kernel_size, self.stride, self.dilation = _pair(kernel_size), _pair(stride), _pair(dilation)

    def get_pad_shape(self, input_shape):
        (ih, iw), (kh, kw), (sh, sw) = input_shape, self.kernel_size, self.stride
        oh, ow = -(-ih // sh), -(-iw // sw)
        return (max((oh - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0),
                max((ow - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0))

    def forward(self, x):
        ph, pw = self.get_pad_shape(x.size()[-2:])
        if ph > 0 or pw > 0:
            if self.padding == "corner":
                x = F.pad(x, [0, pw, 0, ph])
            else:
                x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        return x


class PatchEmbed(nn.Module):
    """maskclip_model.py:330-445, the configuration VisionTransformer builds: Conv2d(k = stride = patch), 'corner' padding."""

    def __init__(self, in_channels=3, embed_dims=768, kernel_size=16, bias=False):
        super().__init__()
        self.embed_dims = embed_dims
        self.adap_padding = AdaptivePadding(kernel_size=kernel_size, stride=kernel_size, padding="corner")
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size=kernel_size, stride=kernel_size, bias=bias)
        self.norm = None

    def forward_rows(self, x):
        rows, hw = _ops.vit_patchify(x, self.projection.kernel_size[0])
        tok = _ops.gemm_tf32_ex(rows, self.projection.weight.reshape(self.embed_dims, -1), self.projection.bias)
        return tok, hw

    def forward(self, x):
        _require(x)
        with torch.no_grad():
            tok, hw = self.forward_rows(x)
        return tok.view(x.shape[0], hw[0] * hw[1], self.embed_dims), hw


def _require(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise OpenESSB200Error("openess_b200 MaskCLIP mirror runs on CUDA tensors only (no CPU fallback)")


class VisionTransformer(nn.Module):
    """maskclip_model.py:545-851 with the reference's defaults (ViT-B/16, pre_norm, final_norm, return_qkv on the last layer)."""

    def __init__(self, img_size=(224, 224), patch_size=16, patch_bias=False, in_channels=3, embed_dims=768, num_layers=12,
                 num_heads=12, mlp_ratio=4, out_indices=-1, qkv_bias=True, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., with_cls_token=True, output_cls_token=False, norm_cfg=dict(type='LN', eps=1e-6),
                 act_cfg=dict(type='GELU'), patch_norm=False, pre_norm=True, final_norm=True, return_qkv=True,
                 skip_last_attn=False, interpolate_mode='bicubic', num_fcs=2, norm_eval=False, with_cp=False,
                 pretrained=None, init_cfg=None):
        super().__init__()
        if drop_rate or attn_drop_rate or drop_path_rate or patch_norm or output_cls_token or not with_cls_token \
                or num_fcs != 2 or act_cfg.get("type") != "GELU" or norm_cfg.get("type") != "LN":
            raise NotImplementedError("only the configuration maskClipFeatureExtractor builds (maskclip_model.py:873) is mirrored")
        if embed_dims % 128 or embed_dims > 1024 or embed_dims != 64 * num_heads:
            raise NotImplementedError("kernels are built for head dim 64 and embed_dims % 128 == 0 (ViT-B/16: 768 = 12 x 64)")
        self.img_size = _pair(img_size)
        self.patch_size = patch_size
        self.interpolate_mode = interpolate_mode
        self.norm_eval, self.with_cp, self.pretrained = norm_eval, with_cp, pretrained
        eps = norm_cfg.get("eps", 1e-5)
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size, bias=patch_bias)
        num_patches = (self.img_size[0] // patch_size) * (self.img_size[1] // patch_size)
        self.with_cls_token, self.output_cls_token = with_cls_token, output_cls_token
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dims))
        self.drop_after_pos = nn.Dropout(p=drop_rate)
        if isinstance(out_indices, int):
            self.out_indices = [num_layers - 1 if out_indices == -1 else out_indices]
        elif isinstance(out_indices, (list, tuple)):
            self.out_indices = out_indices
        else:
            raise TypeError('out_indices must be type of int, list or tuple')
        self.layers = nn.ModuleList(
            TransformerEncoderLayer(embed_dims, num_heads, mlp_ratio * embed_dims, qkv_bias=qkv_bias, eps=eps)
            for _ in range(num_layers))
        self.pre_norm = pre_norm
        if pre_norm:
            self.ln0 = nn.LayerNorm(embed_dims, eps=eps)
        self.final_norm = final_norm
        if final_norm:
            self.ln1 = nn.LayerNorm(embed_dims, eps=eps)
        self.return_qkv = [False] * num_layers
        if isinstance(return_qkv, bool):
            for i in self.out_indices:
                self.return_qkv[i] = return_qkv
        elif isinstance(return_qkv, (list, tuple)):
            for j, i in enumerate(self.out_indices):
                self.return_qkv[i] = return_qkv[j]
        else:
            raise TypeError('return_qkv must be type of bool, list or tuple')
        self.skip_last_attn = skip_last_attn
        self._pos_cache = {}

    @property
    def norm0(self):
        return self.ln0

    @property
    def norm1(self):
        return self.ln1

    @staticmethod
    def resize_pos_embed(pos_embed, input_shpae, pos_shape, mode):
        """:769-797 (mmseg.ops.resize == F.interpolate)."""
        assert pos_embed.ndim == 3, 'shape of pos_embed must be [B, L, C]'
        pos_h, pos_w = pos_shape
        cls_w = pos_embed[:, 0]
        grid = pos_embed[:, (-1 * pos_h * pos_w):].reshape(1, pos_h, pos_w, pos_embed.shape[2]).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=input_shpae, align_corners=False, mode=mode)
        return torch.cat((cls_w.unsqueeze(1), torch.flatten(grid, 2).transpose(1, 2)), dim=1)

    def _pos_for(self, hw, T):
        """Position table [T, D] for an h x w patch grid: parameter preparation, cached per grid / parameter version."""
        key = (hw, self.pos_embed._version, self.pos_embed.data_ptr())
        pos = self._pos_cache.get(key)
        if pos is None:
            pe = self.pos_embed.detach()
            if pe.shape[1] != T:
                ph, pw = self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size
                if pe.shape[1] != ph * pw + 1:
                    raise ValueError('Unexpected shape of pos_embed, got {}.'.format(pe.shape))
                pe = self.resize_pos_embed(pe, hw, (ph, pw), self.interpolate_mode)
            pos = pe.reshape(T, -1).float().contiguous()
            self._pos_cache = {key: pos}
        return pos

    def forward_rows(self, inputs):
        """-> (x, q, k, v, (h, w), T) of the LAST layer, all [B * T, D] row-major with the class token at row 0 of each sample."""
        _require(inputs)
        B = inputs.shape[0]
        tok, hw = self.patch_embed.forward_rows(inputs)
        T = hw[0] * hw[1] + 1
        D = tok.shape[1]
        x = _ops.vit_assemble(tok, self.cls_token.reshape(D), self._pos_for(hw, T), B, T)
        if self.pre_norm:
            x = _ops.layernorm_rows(x, self.ln0.weight, self.ln0.bias, self.ln0.eps)
        outs = {}
        L = len(self.layers)
        for i, layer in enumerate(self.layers):
            last = i == L - 1
            x, q, k, v = layer.forward_rows(x, B, T, self.return_qkv[i] or (last and self.skip_last_attn))
            if last:
                if self.final_norm:
                    x = _ops.layernorm_rows(x, self.ln1.weight, self.ln1.bias, self.ln1.eps)
                    if self.return_qkv[i]:
                        v = _ops.layernorm_rows(v, self.ln1.weight, self.ln1.bias, self.ln1.eps)
                if self.skip_last_attn:
                    x.view(B, T, D)[:, 1:] = v.view(B, T, D)[:, 1:]
            if i in self.out_indices:
                outs[i] = (x if last else x.clone(), q, k, v)
        return outs, hw, T

    def forward(self, inputs):
        """Reference return structure (:808-851): tuple over out_indices of [out NCHW, q [B, hw, C], k, v NCHW] (or out)."""
        with torch.no_grad():
            outs, (h, w), T = self.forward_rows(inputs)
            B = inputs.shape[0]
            res = []
            for i in self.out_indices:
                x, q, k, v = outs[i]
                D = x.shape[1]
                out = x.view(B, T, D)[:, 1:].reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous()
                if self.return_qkv[i]:
                    q = q.view(B, T, D)[:, 1:]
                    k = k.view(B, T, D)[:, 1:]
                    v = v.view(B, T, D)[:, 1:].reshape(B, h, w, D).permute(0, 3, 1, 2).contiguous()
                    out = [out, q, k, v]
                res.append(out)
        return tuple(res)


class MaskClipHead(nn.Module):
    """maskclip_model.py:52-222, the `vit=True` configuration (proj 768 -> 512 without bias, cosine classifier against the
    text embeddings).  `text_embeddings_path=None` -> learnable N(0, 0.01) embeddings as the reference (:104-106);
    `visual_projs_path=None` keeps the random `proj` (the reference would fail in torch.load: used by seeded-weight tests)."""

    def __init__(self, text_embeddings_path='', visual_projs_path='', channels=0, num_classes=16, in_channels=768,
                 in_index=-1, ignore_index=255, align_corners=False, text_categories=16, text_channels=512, vit=True,
                 ks_thresh=1, pd_thresh=0.5, attn_pooling=False, num_heads=32, **kwargs):
        super().__init__()
        if not vit or attn_pooling:
            raise NotImplementedError("only the vit=True head maskClipFeatureExtractor builds is mirrored")
        self.in_channels, self.channels, self.num_classes = in_channels, channels, num_classes
        self.in_index, self.ignore_index, self.align_corners = in_index, ignore_index, align_corners
        self.text_categories, self.text_channels = text_categories, text_channels
        self.text_embeddings_path, self.visual_projs_path = text_embeddings_path, visual_projs_path
        if channels > 0:
            self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
        if text_embeddings_path is None:
            self.text_embeddings = nn.Parameter(torch.zeros(text_categories, text_channels))
            nn.init.normal_(self.text_embeddings, mean=0.0, std=0.01)
        else:
            self.register_buffer('text_embeddings', torch.randn(text_categories, text_channels))
            self.load_text_embeddings()
        self.vit = vit
        self.proj = nn.Conv2d(in_channels, text_channels, 1, bias=False)
        if visual_projs_path is not None:
            self.load_visual_projs()
        self.ks_thresh, self.pd_thresh, self.attn_pooling, self.num_heads = ks_thresh, pd_thresh, attn_pooling, num_heads
        self.image_mapping_local = nn.Conv2d(in_channels, 512, 1)      # constructed, unused by forward (:153)

    def load_text_embeddings(self):
        loaded = torch.load(self.text_embeddings_path, map_location='cpu')
        with torch.no_grad():
            self.text_embeddings[:, :] = loaded[:, :]

    def load_visual_projs(self):
        loaded = torch.load(self.visual_projs_path, map_location='cpu')
        sd = dict(loaded['proj'])
        for key in sd:
            if 'weight' in key and sd[key].ndim == 2:
                sd[key] = sd[key][:, :, None, None]
        self.proj.load_state_dict(sd)

    def logits_rows(self, v_rows):
        """v_rows [n, in_channels] (channels-last pixels) -> cosine logits [n, text_categories] (:177-180, 217-221)."""
        feat = _ops.gemm_tf32_ex(v_rows, self.proj.weight.reshape(self.text_channels, self.in_channels))
        _ops.l2norm_rows_(feat)
        return _ops.gemm_tf32_ex(feat, self.text_embeddings)

    def forward(self, inputs):
        x = inputs[self.in_index]
        v = None
        if isinstance(x, (list, tuple)) and len(x) == 4:
            x, _, _, v = x
        src = v if v is not None else x
        _require(src)
        with torch.no_grad():
            B, C, h, w = src.shape
            rows = src.permute(0, 2, 3, 1).reshape(B * h * w, C).contiguous()
            logits = self.logits_rows(rows).view(B, h, w, -1).permute(0, 3, 1, 2).contiguous()
        return src, logits     # (image_feats, logists), :216 -- image_feats is only defined on the v path in the reference

    def cls_seg(self, feat):
        _require(feat)
        with torch.no_grad():
            B, C, h, w = feat.shape
            rows = feat.permute(0, 2, 3, 1).reshape(B * h * w, C).contiguous().clone()
            _ops.l2norm_rows_(rows)
            return _ops.gemm_tf32_ex(rows, self.text_embeddings).view(B, h, w, -1).permute(0, 3, 1, 2).contiguous()


class maskClipFeatureExtractor(nn.Module):
    """maskclip_model.py:853-915: frozen ViT-B/16 encoder + MaskCLIP head; forward(img [B, 3, H, W]) -> logits [B, K, H, W]
    (bilinear resize of the patch-level cosine logits, align_corners=False).  `maskclip_checkpoint=None` skips the load."""

    def __init__(self, text_embeddings_path, visual_projs_path, text_categories, maskclip_checkpoint, preprocessing=None,
                 test_cfg=dict(mode='whole')):
        super().__init__()
        self.encoder = VisionTransformer()
        self.decoder = MaskClipHead(text_embeddings_path=text_embeddings_path, visual_projs_path=visual_projs_path,
                                    text_categories=text_categories)
        self.align_corners = self.decoder.align_corners
        self.num_classes = self.decoder.num_classes
        self.test_cfg = test_cfg
        self.checkpoint = maskclip_checkpoint
        if maskclip_checkpoint is not None:
            self.encoder = load_checkpoint1(self.checkpoint, self.encoder)
        for p in self.encoder.parameters():
            p.requires_grad = False
        for p in self.decoder.parameters():
            p.requires_grad = False

    def forward(self, img):
        _require(img)
        enc = self.encoder
        if len(enc.out_indices) != 1 or enc.out_indices[0] != len(enc.layers) - 1 or not enc.return_qkv[-1]:
            x = enc(img)                                   # generic structure: go through the reference-shaped outputs
            _, logits = self.decoder(x)
            return F.interpolate(logits, size=img.shape[2:], mode='bilinear', align_corners=self.align_corners)
        with torch.no_grad():
            B = img.shape[0]
            outs, (h, w), T = enc.forward_rows(img)
            v = outs[len(enc.layers) - 1][3]
            D = v.shape[1]
            rows = v.view(B, T, D)[:, 1:].reshape(B * h * w, D)          # drop the class token (:838-842)
            logits = self.decoder.logits_rows(rows)                      # [B * h * w, K] channels-last
            return _ops.bilinear_tokens_to_nchw(logits, B, h, w, (img.shape[2], img.shape[3]))
