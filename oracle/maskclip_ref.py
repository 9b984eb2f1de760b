"""TEST INFRASTRUCTURE ONLY (oracle): plain-torch restatement of the reference's MaskCLIP ViT-B/16 forward
(models/maskclip_model.py).  Only tests/ may import this file; nothing under openess_b200/ does.

PARITY UNPINNED: the reference module cannot be imported here -- its arithmetic lives partly in two un-vendored third
parties, mmcv-full 1.6.0 (mmcv.cnn.bricks.transformer.{MultiheadAttention, FFN}, build_norm_layer, build_conv_layer) and
mmsegmentation 0.30.0 (mmseg.ops.resize), versions from docs/INSTALL.md:154,156 -- and the reference holds no test, golden
vector or trainer call for it.  The restatement follows the reference's own source line by line where it is
in the repo and the published definition of the two mmcv bricks where it is not:
  * mmcv MultiheadAttention(batch_first=True): `out = nn.MultiheadAttention(embed, heads, bias)(q, k, v)[0]` on [L, N, C]
    transposed inputs, returns `identity + dropout_layer(proj_drop(out))` with both rates 0;
  * mmcv FFN(num_fcs=2): `identity + Linear(GELU(Linear(x)))`;
  * build_norm_layer(dict(type='LN', eps=1e-6), C, postfix=n) -> ('ln{n}', nn.LayerNorm(C, eps=1e-6));
  * mmseg.ops.resize(...) == F.interpolate(...).
Module attribute names equal the reference's, so `state_dict()` keys are the reference's keys.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MultiheadAttention(nn.Module):          # mmcv/cnn/bricks/transformer.py (1.6.0), batch_first=True, dropouts 0
    def __init__(self, embed_dims, num_heads, bias=True):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout=0.0, bias=bias)

    def forward(self, query, identity=None):
        if identity is None:
            identity = query
        q = query.transpose(0, 1)
        out = self.attn(query=q, key=q, value=q, need_weights=False)[0].transpose(0, 1)
        return identity + out


class FFN(nn.Module):                         # mmcv/cnn/bricks/transformer.py (1.6.0), num_fcs=2, GELU, dropouts 0
    def __init__(self, embed_dims, feedforward_channels):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.GELU(), nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))

    def forward(self, x, identity=None):
        if identity is None:
            identity = x
        return identity + self.layers(x)


class TransformerEncoderLayer(nn.Module):     # maskclip_model.py:448-541
    def __init__(self, embed_dims, num_heads, feedforward_channels, qkv_bias=True):
        super().__init__()
        self.ln1 = nn.LayerNorm(embed_dims, eps=1e-6)
        self.attn = MultiheadAttention(embed_dims, num_heads, bias=qkv_bias)
        self.ln2 = nn.LayerNorm(embed_dims, eps=1e-6)
        self.ffn = FFN(embed_dims, feedforward_channels)

    def forward(self, x, return_qkv=False):   # :519-541
        q, k, v = None, None, None
        if return_qkv:
            y = self.ln1(x)
            y = F.linear(y, self.attn.attn.in_proj_weight, self.attn.attn.in_proj_bias)
            N, L, C = y.shape
            y = y.view(N, L, 3, C // 3).permute(2, 0, 1, 3).reshape(3 * N, L, C // 3)
            y = F.linear(y, self.attn.attn.out_proj.weight, self.attn.attn.out_proj.bias)
            nn_ = y.shape[0]
            q, k, v = y[:nn_ // 3], y[nn_ // 3:(nn_ // 3) * 2], y[(nn_ // 3) * 2:]
            v = v + x                          # the reference writes `v += x` on a view of y; same values
            v = self.ffn(self.ln2(v), identity=v)
        x = self.attn(self.ln1(x), identity=x)
        x = self.ffn(self.ln2(x), identity=x)
        return x, q, k, v


class PatchEmbed(nn.Module):                  # maskclip_model.py:330-445 (Conv2d k = stride = patch, 'corner' padding)
    def __init__(self, in_channels, embed_dims, patch, bias):
        super().__init__()
        self.patch = patch
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size=patch, stride=patch, bias=bias)

    def forward(self, x):
        H, W = x.shape[-2:]
        ph = (-H) % self.patch                 # AdaptivePadding.get_pad_shape :296-307 with kernel = stride, dilation 1
        pw = (-W) % self.patch
        if ph or pw:
            x = F.pad(x, [0, pw, 0, ph])
        x = self.projection(x)
        hw = (x.shape[2], x.shape[3])
        return x.flatten(2).transpose(1, 2), hw


class VisionTransformer(nn.Module):           # maskclip_model.py:545-851, the defaults maskClipFeatureExtractor uses
    def __init__(self, img_size=(224, 224), patch_size=16, in_channels=3, embed_dims=768, num_layers=12, num_heads=12,
                 mlp_ratio=4, patch_bias=False, skip_last_attn=False):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size, patch_bias)
        n = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dims))
        self.layers = nn.ModuleList(TransformerEncoderLayer(embed_dims, num_heads, mlp_ratio * embed_dims)
                                    for _ in range(num_layers))
        self.ln0 = nn.LayerNorm(embed_dims, eps=1e-6)     # pre_norm
        self.ln1 = nn.LayerNorm(embed_dims, eps=1e-6)     # final_norm
        self.skip_last_attn = skip_last_attn

    def _pos(self, x, hw):                    # :738-797
        pe = self.pos_embed
        if x.shape[1] != pe.shape[1]:
            ph, pw = self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size
            assert pe.shape[1] == ph * pw + 1
            grid = pe[:, -ph * pw:].reshape(1, ph, pw, pe.shape[2]).permute(0, 3, 1, 2)
            grid = F.interpolate(grid, size=hw, align_corners=False, mode='bicubic')
            pe = torch.cat((pe[:, 0].unsqueeze(1), torch.flatten(grid, 2).transpose(1, 2)), dim=1)
        return x + pe

    def forward(self, inputs):                # :799-851 with out_indices = [last], return_qkv = True on the last layer
        B = inputs.shape[0]
        x, hw = self.patch_embed(inputs)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        x = self.ln0(self._pos(x, hw))
        L = len(self.layers)
        for i, layer in enumerate(self.layers):
            x, q, k, v = layer(x, i == L - 1)
            if i == L - 1:
                x = self.ln1(x)
                v = self.ln1(v)
                if self.skip_last_attn:
                    x = torch.cat((x[:, :1], v[:, 1:]), dim=1)
        C = x.shape[2]
        out = x[:, 1:].reshape(B, hw[0], hw[1], C).permute(0, 3, 1, 2).contiguous()
        q, k = q[:, 1:], k[:, 1:]
        v = v[:, 1:].reshape(B, hw[0], hw[1], C).permute(0, 3, 1, 2).contiguous()
        return ([out, q, k, v],)


class MaskClipHead(nn.Module):                # maskclip_model.py:52-222, vit=True
    def __init__(self, text_categories, in_channels=768, text_channels=512):
        super().__init__()
        self.text_embeddings = nn.Parameter(torch.zeros(text_categories, text_channels))
        nn.init.normal_(self.text_embeddings, mean=0.0, std=0.01)
        self.proj = nn.Conv2d(in_channels, text_channels, 1, bias=False)
        self.image_mapping_local = nn.Conv2d(in_channels, 512, 1)

    def forward(self, inputs):                # :157-221
        x, q, k, v = inputs[-1]
        feat = self.proj(v)
        feat = feat / feat.norm(dim=1, keepdim=True)
        return v, F.conv2d(feat, self.text_embeddings[:, :, None, None])


class maskClipFeatureExtractor(nn.Module):    # maskclip_model.py:853-915
    def __init__(self, text_categories, **vit_kwargs):
        super().__init__()
        self.encoder = VisionTransformer(**vit_kwargs)
        self.decoder = MaskClipHead(text_categories, in_channels=self.encoder.ln0.normalized_shape[0])

    def forward(self, img):
        _, logits = self.decoder(self.encoder(img))
        return F.interpolate(logits, size=(img.shape[2], img.shape[3]), mode='bilinear', align_corners=False)


def seed_weights(model, seed=1205):
    """Deterministic non-degenerate weights (the reference ships none): N(0, 0.02)-style init of a CLIP checkpoint's scale."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("ln0.weight") or name.endswith("ln1.weight") or name.endswith("ln2.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.ndim == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "text_embeddings" in name:
                p.copy_(torch.randn(p.shape, generator=g))
            elif "pos_embed" in name or "cls_token" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / fan_in ** 0.5))
    return model
