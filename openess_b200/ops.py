"""Small dense ops over the C ABI with autograd: the per-pixel linear map used by the fused segmentation head."""
import ctypes
import os

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


def _pixel_linear_raw(x, W, bias):
    B, Cin = x.shape[:2]
    Cout = W.shape[0]
    HW = x.numel() // (B * Cin)
    y = torch.empty((B, Cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_pixel_linear(ptr(x), ptr(W), ptr(bias), B, Cin, Cout, HW, ptr(y), stream_ptr(x.device)),
              "oess_pixel_linear")
    return y


class _PixelLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, bias):
        require_cuda(x, W)
        x, W = _f32c(x), _f32c(W)
        b = None if bias is None else _f32c(bias)
        ctx.save_for_backward(x, W)
        ctx.has_bias = bias is not None
        return _pixel_linear_raw(x, W, b)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = _f32c(dy)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = _pixel_linear_raw(dy, W.t().contiguous(), None)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            B, Cin = x.shape[:2]
            Cout = W.shape[0]
            HW = x.numel() // (B * Cin)
            nb = ctypes.c_size_t(0)
            check(lib().oess_pixel_linear_wgrad_ws_bytes(Cin, Cout, ctypes.byref(nb)), "oess_pixel_linear_wgrad_ws_bytes")
            dW = torch.empty_like(W)
            db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            with torch.cuda.device(x.device):
                ws = _lib.workspace(nb.value, x.device)
                check(lib().oess_pixel_linear_wgrad(ptr(dy), ptr(x), B, Cin, Cout, HW, ptr(dW), ptr(db), ptr(ws),
                                                    ws.numel(), stream_ptr(x.device)), "oess_pixel_linear_wgrad")
        return dx, dW, db


def pixel_linear(x, W, bias=None):
    """y[b, k, ...] = bias[k] + sum_c W[k, c] x[b, c, ...]  for x [B, Cin, *spatial], Cin and Cout <= 64."""
    if W.ndim != 2 or W.shape[1] != x.shape[1]:
        raise ValueError("W must be [Cout, Cin]")
    if W.shape[0] > 64 or W.shape[1] > 64:
        raise ValueError("pixel_linear supports at most 64 input and 64 output channels")
    return _PixelLinear.apply(x, W, bias)


def gemm_tf32(a, b, bias=None, out=None):
    """out[M, N] = a[M, K] @ b[N, K].T + bias  on the tensor cores (tcgen05.mma.kind::tf32, fp32 accumulate).

    a, b: contiguous float32 CUDA tensors (b in nn.Linear weight layout); K % 4 == 0."""
    _lib.require_cuda(a, b, bias)
    a, b = _f32c(a), _f32c(b)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("gemm_tf32: a must be [M, K] and b [N, K]")
    M, K = a.shape
    N = b.shape[0]
    if bias is not None:
        bias = _f32c(bias)
        if bias.numel() != N:
            raise ValueError("gemm_tf32: bias must have N elements")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().oess_gemm_tf32(ptr(a), ptr(b), ptr(bias), ptr(out), M, N, K, stream_ptr(a.device)), "oess_gemm_tf32")
    return out


def gemm_tf32_ex(a, b, bias=None, residual=None, act=None, out=None, round_out=False):
    """out = act(a @ b.T + bias) + residual on the tensor cores (oess_gemm_tf32_ex): act in (None, "gelu"); `out` may be
    `residual` (the residual add of a transformer block, maskclip_model.py:538-539)."""
    _lib.require_cuda(a, b, bias, residual)
    a, b = _f32c(a), _f32c(b)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("gemm_tf32_ex: a must be [M, K] and b [N, K]")
    M, K = a.shape
    N = b.shape[0]
    if bias is not None:
        bias = _f32c(bias)
        if bias.numel() != N:
            raise ValueError("gemm_tf32_ex: bias must have N elements")
    if residual is not None:
        if residual.shape != (M, N) or residual.dtype != torch.float32 or not residual.is_contiguous():
            raise ValueError("gemm_tf32_ex: residual must be a contiguous float32 [M, N] tensor")
    if act not in (None, "gelu"):
        raise ValueError("gemm_tf32_ex: act must be None or 'gelu'")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().oess_gemm_tf32_ex(ptr(a), ptr(b), ptr(bias), ptr(residual), ptr(out), M, N, K,
                                      (1 if act == "gelu" else 0) | (2 if round_out else 0), stream_ptr(a.device)),
              "oess_gemm_tf32_ex")
    return out


def vit_patchify(img, patch):
    """[B, C, H, W] -> [B * ceil(H / P) * ceil(W / P), C * P * P] rows in projection.weight.view(D, -1) column order, zero
    'corner' padding (maskclip_model.py:303-312, 427-441)."""
    _lib.require_cuda(img)
    img = _f32c(img)
    B, C, H, W = img.shape
    h, w = -(-H // patch), -(-W // patch)
    rows = torch.empty(B * h * w, C * patch * patch, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        check(lib().oess_vit_patchify(ptr(img), B, C, H, W, patch, ptr(rows), stream_ptr(img.device)), "oess_vit_patchify")
    return rows, (h, w)


def vit_assemble(tok, cls, pos, B, T):
    """x[b, 0] = cls + pos[0], x[b, 1 + i] = tok[b, i] + pos[1 + i]  ->  [B * T, D] (maskclip_model.py:799-806)."""
    _lib.require_cuda(tok, cls, pos)
    tok, cls, pos = _f32c(tok), _f32c(cls), _f32c(pos)
    D = cls.numel()
    if tok.shape != (B * (T - 1), D) or pos.numel() != T * D:
        raise ValueError("vit_assemble: tok must be [B * (T - 1), D] and pos [T, D]")
    x = torch.empty(B * T, D, dtype=torch.float32, device=tok.device)
    with torch.cuda.device(tok.device):
        check(lib().oess_vit_assemble(ptr(tok), ptr(cls), ptr(pos), B, T, D, ptr(x), stream_ptr(tok.device)), "oess_vit_assemble")
    return x


def layernorm_rows(x, weight, bias, eps):
    """nn.LayerNorm over the last dim of a [rows, D] tensor (D % 128 == 0, D <= 1024)."""
    _lib.require_cuda(x, weight, bias)
    x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
    rows, D = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().oess_layernorm_rows(ptr(x), ptr(weight), ptr(bias), float(eps), rows, D, ptr(y), stream_ptr(x.device)),
              "oess_layernorm_rows")
    return y


MHA_TENSOR_CORES = os.environ.get("OESS_MHA", "tc") != "simt"


def mha_fwd(qkv, B, T, heads, tensor_cores=None):
    """softmax(q k^T / 8) v per head on the packed in_proj output [B * T, 3 * heads * 64] -> [B * T, heads * 64].
    tensor_cores (default: module switch MHA_TENSOR_CORES / env OESS_MHA=simt|tc): oess_mha_fwd_tc (tcgen05, TF32 operands)
    or oess_mha_fwd (exact fp32 on the FMA pipes)."""
    _lib.require_cuda(qkv)
    qkv = _f32c(qkv)
    D = heads * 64
    if qkv.shape != (B * T, 3 * D):
        raise ValueError("mha_fwd: qkv must be [B * T, 3 * heads * 64] (head dim 64)")
    out = torch.empty(B * T, D, dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        fn = lib().oess_mha_fwd_tc if (MHA_TENSOR_CORES if tensor_cores is None else tensor_cores) else lib().oess_mha_fwd
        check(fn(ptr(qkv), B, T, heads, ptr(out), stream_ptr(qkv.device)), "oess_mha_fwd")
    return out


def l2norm_rows_(x):
    """x /= ||x||_2 per row, in place (maskclip_model.py:218-219)."""
    _lib.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.ndim != 2:
        raise ValueError("l2norm_rows_: contiguous float32 [rows, D]")
    with torch.cuda.device(x.device):
        check(lib().oess_l2norm_rows(ptr(x), x.shape[0], x.shape[1], stream_ptr(x.device)), "oess_l2norm_rows")
    return x


def bilinear_tokens_to_nchw(tok, B, h, w, size):
    """F.interpolate(bilinear, align_corners=False) of a channels-last token map [B * h * w, K] -> [B, K, H, W]."""
    _lib.require_cuda(tok)
    tok = _f32c(tok)
    K = tok.shape[1]
    if tok.shape[0] != B * h * w:
        raise ValueError("bilinear_tokens_to_nchw: tok must be [B * h * w, K]")
    H, W = size
    out = torch.empty(B, K, H, W, dtype=torch.float32, device=tok.device)
    with torch.cuda.device(tok.device):
        check(lib().oess_bilinear_tokens_to_nchw(ptr(tok), B, h, w, K, H, W, ptr(out), stream_ptr(tok.device)),
              "oess_bilinear_tokens_to_nchw")
    return out


def convlstm_pack(weight, bias, hidden):
    """Repack ConvLSTM.Gates (e2vid/model/submodules.py:186: Conv2d(2C, 4C, 3, padding=1), input = cat(x, h)) for
    oess_convlstm_step_nhwc: rows (chunk, gate, c) so one 256-column tile holds all four gates of 64 hidden channels,
    columns (source, tap, channel) in the kernel's K order."""
    C = int(hidden)
    if weight.shape != (4 * C, 2 * C, 3, 3) or C % 64:
        raise ValueError("convlstm_pack: expects Gates.weight [4C, 2C, 3, 3] with input_size == hidden_size, C % 64 == 0")
    w = weight.detach().float().reshape(4, C // 64, 64, 2, C, 9)           # gate, chunk, c, source, ch, tap
    w = w.permute(1, 0, 2, 3, 5, 4).contiguous().reshape(4 * C, 2 * 9 * C)  # chunk, gate, c | source, tap, ch
    b = bias.detach().float().reshape(4, C // 64, 64).permute(1, 0, 2).contiguous().reshape(4 * C)
    return w, b


def convlstm_step(x, prev_state, w_packed, b_packed):
    """(hidden, cell) of one ConvLSTM step on the tensor cores.  x: [B, C, H, W]; prev_state None or (h, c).
    Tensors are handled channels-last (the kernel's layout); the returned tensors are [B, C, H, W] channels-last."""
    _lib.require_cuda(x, w_packed, b_packed)
    B, C, H, W = x.shape
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    hp = cp = None
    if prev_state is not None:
        hp = prev_state[0].float().contiguous(memory_format=cl)
        cp = prev_state[1].float().contiguous(memory_format=cl)
    h = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device, memory_format=cl)
    c = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device, memory_format=cl)
    with torch.cuda.device(x.device):
        check(lib().oess_convlstm_step_nhwc(ptr(xc), ptr(hp), ptr(cp), ptr(w_packed), ptr(b_packed), ptr(h), ptr(c),
                                            B, H, W, C, stream_ptr(x.device)), "oess_convlstm_step_nhwc")
    return h, c


def convlstm_step_bf16(x_bf, prev_state, w_packed_bf, b_packed):
    """One ConvLSTM step with bf16 tensor-core operands (oess_convlstm_step_nhwc_bf16; frozen E2VID only).
    x_bf: [B, C, H, W] bfloat16 channels-last; prev_state None or (h_fp32, c_fp32) where h carries its bf16 copy as the
    attribute `_oess_bf16` (set by this function; converted on the fly otherwise); w_packed_bf: convlstm_pack(...)[0] in
    bfloat16.  Returns (h, c) fp32 channels-last, h with `_oess_bf16` attached."""
    _lib.require_cuda(x_bf, w_packed_bf, b_packed)
    B, C, H, W = x_bf.shape
    cl = torch.channels_last
    if x_bf.dtype != torch.bfloat16 or not x_bf.is_contiguous(memory_format=cl):
        x_bf = x_bf.to(torch.bfloat16).contiguous(memory_format=cl)
    hp = cp = None
    if prev_state is not None:
        hp = getattr(prev_state[0], "_oess_bf16", None)
        if hp is None:
            hp = prev_state[0].to(torch.bfloat16).contiguous(memory_format=cl)
        cp = prev_state[1].float().contiguous(memory_format=cl)
    h = torch.empty((B, C, H, W), dtype=torch.float32, device=x_bf.device, memory_format=cl)
    hb = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=x_bf.device, memory_format=cl)
    c = torch.empty((B, C, H, W), dtype=torch.float32, device=x_bf.device, memory_format=cl)
    with torch.cuda.device(x_bf.device):
        check(lib().oess_convlstm_step_nhwc_bf16(ptr(x_bf), ptr(hp), ptr(cp), ptr(w_packed_bf), ptr(b_packed), ptr(h), ptr(hb),
                                                 ptr(c), B, H, W, C, stream_ptr(x_bf.device)), "oess_convlstm_step_nhwc_bf16")
    h._oess_bf16 = hb
    return h, c


def conv2d_tc_bf16out(x, w_packed, bias, kernel_size, stride=1, padding=0, dilation=1, relu=False):
    """conv2d_tc whose output is stored ONLY as bfloat16 (channels-last): the x operand of convlstm_step_bf16."""
    _lib.require_cuda(x, w_packed, bias)
    B, Cin, H, W = x.shape
    KH, KW = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    Cout = w_packed.shape[0]
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    yb = torch.empty((B, Cout, Ho, Wo), dtype=torch.bfloat16, device=x.device, memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        check(lib().oess_conv2d_nhwc_tf32_bf16out(ptr(xc), ptr(w_packed), ptr(bc), None, ptr(yb), B, H, W, Cin, Cout, KH, KW,
                                                  stride, padding, dilation, 1 if relu else 0, stream_ptr(x.device)),
              "oess_conv2d_nhwc_tf32_bf16out")
    return yb


def conv2d_pack(weight):
    """[Cout, Cin, KH, KW] -> [Cout, KH * KW * Cin_p] (Cin_p = Cin rounded up to 32, zero padded), column (tap, channel):
    the K order of oess_conv2d_nhwc_tf32."""
    Cout, Cin, KH, KW = weight.shape
    cin_p = (Cin + 31) // 32 * 32
    w = torch.zeros(Cout, KH * KW, cin_p, dtype=torch.float32, device=weight.device)
    w[:, :, :Cin] = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin)
    return round_tf32(w.reshape(Cout, KH * KW * cin_p).contiguous())


def round_tf32(t):
    """fp32 -> nearest TF32 value (ties away from zero, = cvt.rna.tf32.f32), kept as fp32.  tcgen05.mma.kind::tf32 truncates
    its operands; rounding the packed weights once removes the weight half of that systematic bias."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32)


def conv2d_tc(x, w_packed, bias, kernel_size, stride=1, padding=0, dilation=1, relu=False, residual=None, round_out=False):
    """Convolution on the tensor cores (tcgen05 implicit GEMM, TF32 operands, fp32 accumulate).
    x: [B, Cin, H, W] (any memory format; handled channels-last), Cin % 4 == 0; returns [B, Cout, Ho, Wo] channels-last.
    round_out: store the output rounded to TF32 (for activations that feed the next tensor-core conv, see round_tf32)."""
    _lib.require_cuda(x, w_packed, bias, residual)
    B, Cin, H, W = x.shape
    KH, KW = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    Cout = w_packed.shape[0]
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=cl)
    rc = None if residual is None else residual.float().contiguous(memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        check(lib().oess_conv2d_nhwc_tf32(ptr(xc), ptr(w_packed), ptr(bc), ptr(rc), ptr(y), B, H, W, Cin, Cout, KH, KW,
                                          stride, padding, dilation, (1 if relu else 0) | (2 if round_out else 0),
                                          stream_ptr(x.device)), "oess_conv2d_nhwc_tf32")
    return y


def _bn_momentum(bn):
    """exponential_average_factor of torch.nn.BatchNorm2d: `momentum`, or the cumulative moving average 1 / num_batches_tracked
    (counted AFTER this batch) when momentum is None (torch/nn/modules/batchnorm.py)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    nbt = bn.num_batches_tracked
    return 1.0 / (float(nbt) + 1.0) if nbt is not None else 0.0


def batchnorm_nhwc_(x, bn, residual=None, relu=False):
    """In-place torch.nn.BatchNorm2d (module `bn`: batch statistics + running-stat update when bn.training, running
    statistics otherwise) on a channels-last [B, C, H, W] tensor, with optional residual add and ReLU fused."""
    _lib.require_cuda(x)
    B, C, H, W = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last) or x.dtype != torch.float32:
        raise ValueError("batchnorm_nhwc_: x must be a float32 channels-last tensor")
    res = None if residual is None else residual.float().contiguous(memory_format=torch.channels_last)
    training = bn.training or bn.running_mean is None
    mom = _bn_momentum(bn)
    nb = ctypes.c_size_t(0)
    check(lib().oess_bn_ws_bytes(C, ctypes.byref(nb)), "oess_bn_ws_bytes")
    track = training and bn.track_running_stats and bn.running_mean is not None
    with torch.cuda.device(x.device):
        ws = _lib.workspace(nb.value, x.device)
        check(lib().oess_batchnorm_nhwc(ptr(x), B * H * W, C, ptr(bn.weight), ptr(bn.bias),
                                        ptr(bn.running_mean) if (track or not training) else None,
                                        ptr(bn.running_var) if (track or not training) else None,
                                        float(bn.eps), mom, 1 if training else 0, ptr(res), 1 if relu else 0,
                                        ptr(ws), ws.numel(), stream_ptr(x.device)), "oess_batchnorm_nhwc")
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return x


class _UpNormPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, seg, S, M, H, W):
        B, C, h, w = d.shape
        dc = d.float().contiguous(memory_format=torch.channels_last)
        seg = seg.contiguous()
        pooled = torch.empty(M, C, dtype=torch.float32, device=d.device)
        counts = torch.empty(M, dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            check(lib().oess_upnorm_pool_fwd(ptr(dc), ptr(seg), B, h, w, C, H, W, S, M, ptr(pooled), ptr(counts), None,
                                             stream_ptr(d.device)), "oess_upnorm_pool_fwd")
        ctx.save_for_backward(dc, seg)
        ctx.meta = (S, M, H, W)
        ctx.mark_non_differentiable(counts)
        return pooled, counts

    @staticmethod
    def backward(ctx, g_pooled, _g_counts):
        dc, seg = ctx.saved_tensors
        S, M, H, W = ctx.meta
        B, C, h, w = dc.shape
        g = _f32c(g_pooled)
        d_grad = torch.empty_like(dc)                       # channels-last
        with torch.cuda.device(dc.device):
            check(lib().oess_upnorm_pool_bwd(ptr(dc), ptr(seg), ptr(g), B, h, w, C, H, W, S, M, ptr(d_grad),
                                             stream_ptr(dc.device)), "oess_upnorm_pool_bwd")
        return d_grad, None, None, None, None, None


def upnorm_pool(d, superpixels, superpixel_size, M, scale=4):
    """q [M, 256] = superpixel mean-pool (sum / (count + 1e-6), pretrain_trainer.py:461-463) of
    F.normalize(Upsample(scale, bilinear, align_corners=True)(d), dim=1) (image_model.py:121-124,139-141) without ever
    materialising the full-resolution map.  d: [B, 256, h, w]; superpixels: int64 [B, h * scale, w * scale]."""
    _lib.require_cuda(d, superpixels)
    B, C, h, w = d.shape
    H, W = h * scale, w * scale
    if tuple(superpixels.shape) != (B, H, W) or superpixels.dtype != torch.int64:
        raise ValueError("upnorm_pool: superpixels must be int64 [B, h * scale, w * scale]")
    pooled, counts = _UpNormPool.apply(d, superpixels, int(superpixel_size), int(M), H, W)
    return pooled / (counts[:, None] + 1e-6)


def conv_bn_train(x, w_packed, bias, kernel_size, stride, padding, dilation, bn, residual=None, relu=False):
    """act(BatchNorm_train(conv(x)) + residual) with the batch statistics accumulated in the conv's TMEM epilogue
    (oess_conv2d_nhwc_tf32_stats -> oess_batchnorm_nhwc_sums): one pass over the activation less than conv + BN."""
    _lib.require_cuda(x, w_packed)
    B, Cin, H, W = x.shape
    KH = KW = int(kernel_size)
    Cout = w_packed.shape[0]
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=cl)
    res = None if residual is None else residual.float().contiguous(memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    mom = _bn_momentum(bn)
    track = bn.track_running_stats and bn.running_mean is not None
    nb = ctypes.c_size_t(0)
    check(lib().oess_bn_ws_bytes(Cout, ctypes.byref(nb)), "oess_bn_ws_bytes")
    with torch.cuda.device(x.device):
        ws = _lib.workspace(nb.value, x.device)
        st = stream_ptr(x.device)
        check(lib().oess_conv2d_nhwc_tf32_stats(ptr(xc), ptr(w_packed), ptr(bc), ptr(y), B, H, W, Cin, Cout, KH, KW, stride,
                                                padding, dilation, ptr(ws), st), "oess_conv2d_nhwc_tf32_stats")
        check(lib().oess_batchnorm_nhwc_sums(ptr(y), B * Ho * Wo, Cout, ptr(bn.weight), ptr(bn.bias),
                                             ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None,
                                             float(bn.eps), mom, ptr(res), 1 if relu else 0, ptr(ws), ws.numel(), st),
              "oess_batchnorm_nhwc_sums")
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return y


def conv2d_pack_bf16(weight):
    """[Cout, Cin, KH, KW] -> bf16 [Cout, KH * KW * Cin_p] (Cin_p = Cin rounded up to 64, zero padded), column (tap, channel):
    the K order of oess_conv2d_nhwc_bf16."""
    Cout, Cin, KH, KW = weight.shape
    cin_p = (Cin + 63) // 64 * 64
    w = torch.zeros(Cout, KH * KW, cin_p, dtype=torch.float32, device=weight.device)
    w[:, :, :Cin] = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin)
    return w.reshape(Cout, KH * KW * cin_p).to(torch.bfloat16).contiguous()


def _check_bf16_nhwc(x):
    if x.dtype != torch.bfloat16 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("expected a bfloat16 channels-last [B, C, H, W] tensor")


def conv2d_tc_bf16(x_bf, w_packed_bf, bias, kernel_size, stride=1, padding=0, dilation=1, relu=False, residual=None,
                   want_f32=True, want_bf16=True):
    """Frozen-network convolution with bfloat16 operands (tcgen05 kind::f16, fp32 accumulate; oess_conv2d_nhwc_bf16).
    x_bf: bf16 channels-last [B, Cin, H, W], Cin % 8 == 0; residual fp32.  Returns (y fp32 or None, y bf16 or None)."""
    _lib.require_cuda(x_bf, w_packed_bf, bias, residual)
    _check_bf16_nhwc(x_bf)
    B, Cin, H, W = x_bf.shape
    KH = KW = int(kernel_size)
    Cout = w_packed_bf.shape[0]
    cl = torch.channels_last
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x_bf.device, memory_format=cl) if want_f32 else None
    yb = torch.empty((B, Cout, Ho, Wo), dtype=torch.bfloat16, device=x_bf.device, memory_format=cl) if want_bf16 else None
    rc = None if residual is None else residual.float().contiguous(memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    with torch.cuda.device(x_bf.device):
        check(lib().oess_conv2d_nhwc_bf16(ptr(x_bf), ptr(w_packed_bf), ptr(bc), ptr(rc), ptr(y), ptr(yb), B, H, W, Cin, Cout,
                                          KH, KW, stride, padding, dilation, 1 if relu else 0, None, stream_ptr(x_bf.device)),
              "oess_conv2d_nhwc_bf16")
    return y, yb


def conv_bn_train_bf16(x_bf, w_packed_bf, bias, kernel_size, stride, padding, dilation, bn, residual=None, relu=False,
                       want_f32=True):
    """conv_bn_train with bfloat16 conv operands: batch statistics from the fp32 accumulators in the conv epilogue, the
    normalised result stored as bf16 (next conv's operand) and, when want_f32, as fp32 (a later residual / the output).
    Returns (y fp32 or None, y bf16)."""
    _lib.require_cuda(x_bf, w_packed_bf)
    _check_bf16_nhwc(x_bf)
    B, Cin, H, W = x_bf.shape
    KH = KW = int(kernel_size)
    Cout = w_packed_bf.shape[0]
    cl = torch.channels_last
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x_bf.device, memory_format=cl)
    yb = torch.empty((B, Cout, Ho, Wo), dtype=torch.bfloat16, device=x_bf.device, memory_format=cl)
    res = None if residual is None else residual.float().contiguous(memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    mom = _bn_momentum(bn)
    track = bn.track_running_stats and bn.running_mean is not None
    nb = ctypes.c_size_t(0)
    check(lib().oess_bn_ws_bytes(Cout, ctypes.byref(nb)), "oess_bn_ws_bytes")
    with torch.cuda.device(x_bf.device):
        ws = _lib.workspace(nb.value, x_bf.device)
        st = stream_ptr(x_bf.device)
        check(lib().oess_conv2d_nhwc_bf16(ptr(x_bf), ptr(w_packed_bf), ptr(bc), None, ptr(y), None, B, H, W, Cin, Cout, KH, KW,
                                          stride, padding, dilation, 0, ptr(ws), st), "oess_conv2d_nhwc_bf16")
        check(lib().oess_batchnorm_nhwc_sums_bf16(ptr(y), B * Ho * Wo, Cout, ptr(bn.weight), ptr(bn.bias),
                                                  ptr(bn.running_mean) if track else None,
                                                  ptr(bn.running_var) if track else None, float(bn.eps), mom, ptr(res),
                                                  1 if relu else 0, ptr(yb), 1 if want_f32 else 0, ptr(ws), ws.numel(), st),
              "oess_batchnorm_nhwc_sums_bf16")
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return (y if want_f32 else None), yb


class _BilinearResize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        B, C, h, w = x.shape
        xc = x.float().contiguous()
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().oess_bilinear_resize_planes(ptr(xc), B * C, h, w, H, W, ptr(out), stream_ptr(x.device)),
                  "oess_bilinear_resize_planes")
        ctx.dims = (B, C, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, h, w, H, W = ctx.dims
        gc = g.float().contiguous()
        tmp = torch.empty((B * C, H, w), dtype=torch.float32, device=g.device)
        dx = torch.empty((B, C, h, w), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().oess_bilinear_resize_planes_bwd(ptr(gc), B * C, h, w, H, W, ptr(tmp), ptr(dx), stream_ptr(g.device)),
                  "oess_bilinear_resize_planes_bwd")
        return dx, None, None


def bilinear_resize(x, size):
    """F.interpolate(x, size=size, mode='bilinear', align_corners=False) for a CUDA [B, C, h, w] tensor, differentiable: the
    backward is a separable gather instead of torch's atomic scatter (oess_bilinear_resize_planes / _bwd)."""
    _lib.require_cuda(x)
    H, W = int(size[0]), int(size[1])
    return _BilinearResize.apply(x, H, W)


class _Up2xCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip):
        B, C1, H, W = x.shape
        C2 = 0 if skip is None else skip.shape[1]
        cl = torch.channels_last
        xc = x.float().contiguous(memory_format=cl)
        sc = None if skip is None else skip.float().contiguous(memory_format=cl)
        out = torch.empty((B, C1 + C2, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=cl)
        with torch.cuda.device(x.device):
            check(lib().oess_upsample2x_cat_nhwc(ptr(xc), ptr(sc), B, H, W, C1, C2, ptr(out), stream_ptr(x.device)),
                  "oess_upsample2x_cat_nhwc")
        ctx.dims = (B, H, W, C1, C2)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, C1, C2 = ctx.dims
        cl = torch.channels_last
        gc = g.float().contiguous(memory_format=cl)
        need_x, need_s = ctx.needs_input_grad[0], C2 > 0 and ctx.needs_input_grad[1]
        dx = torch.empty((B, C1, H, W), dtype=torch.float32, device=g.device, memory_format=cl) if need_x else None
        ds = torch.empty((B, C2, 2 * H, 2 * W), dtype=torch.float32, device=g.device, memory_format=cl) if need_s else None
        if need_x or need_s:
            with torch.cuda.device(g.device):
                check(lib().oess_upsample2x_cat_nhwc_bwd(ptr(gc), B, H, W, C1, C2, ptr(dx), ptr(ds), stream_ptr(g.device)),
                      "oess_upsample2x_cat_nhwc_bwd")
        return dx, ds


def upsample2x_cat(x, skip=None):
    """cat([F.interpolate(x, scale_factor=2, mode='nearest'), skip], dim=1) in one pass, channels-last, differentiable
    (oess_upsample2x_cat_nhwc; models/style_networks.py:148-158).  skip None: plain nearest x2 upsampling."""
    _lib.require_cuda(x, skip)
    if x.shape[1] % 4 or (skip is not None and (skip.shape[1] % 4 or skip.shape[2:] != (2 * x.shape[2], 2 * x.shape[3]))):
        raise ValueError("upsample2x_cat: channels % 4 == 0 and skip at twice the resolution of x")
    return _Up2xCat.apply(x, skip)


def zero_insert2x_nhwc(x, skip=None):
    """[B, C, H, W] (channels-last) -> [B, C, 2H, 2W] with (x + skip) at the even positions and zeros elsewhere: the input of
    a stride-2 transposed convolution run as a stride-1 convolution (oess_zero_insert2x_nhwc)."""
    _lib.require_cuda(x, skip)
    B, C, H, W = x.shape
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    sc = None if skip is None else skip.float().contiguous(memory_format=cl)
    z = torch.empty((B, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=cl)
    with torch.cuda.device(x.device):
        check(lib().oess_zero_insert2x_nhwc(ptr(xc), ptr(sc), B, H, W, C, ptr(z), stream_ptr(x.device)), "oess_zero_insert2x_nhwc")
    return z


def conv_transpose2x_pack(weight, scale=None):
    """nn.ConvTranspose2d weight [Cin, Cout, K, K] (stride 2) -> packed weights of the equivalent stride-1 convolution over the
    zero-inserted input: W'[o, i, ky, kx] = W[i, o, K-1-ky, K-1-kx] (optionally scaled per output channel: folded BatchNorm)."""
    wt = weight.detach().float().flip(2, 3).permute(1, 0, 2, 3)
    if scale is not None:
        wt = wt * scale.detach().float()[:, None, None, None]
    return conv2d_pack(wt.contiguous())


def pred_sigmoid_nhwc(x, skip, w, bias):
    """sigmoid(1x1 conv to ONE channel of (x + skip)) -> [B, 1, H, W]; x, skip channels-last [B, C, H, W], w [C], bias float."""
    _lib.require_cuda(x, skip, w)
    B, C, H, W = x.shape
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    sc = None if skip is None else skip.float().contiguous(memory_format=cl)
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_pred_sigmoid_nhwc(ptr(xc), ptr(sc), ptr(_f32c(w)), float(bias), B * H * W, C, ptr(out), stream_ptr(x.device)),
              "oess_pred_sigmoid_nhwc")
    return out


def maxpool3x3s2_nhwc(x):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (models/_resnet.py:137) on a channels-last [B, C, H, W] tensor."""
    _lib.require_cuda(x)
    B, C, H, W = x.shape
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    y = torch.empty((B, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=torch.float32, device=x.device, memory_format=cl)
    with torch.cuda.device(x.device):
        check(lib().oess_maxpool3x3s2_nhwc(ptr(xc), B, H, W, C, ptr(y), stream_ptr(x.device)), "oess_maxpool3x3s2_nhwc")
    return y


def global_avgpool_nhwc(x):
    """nn.AdaptiveAvgPool2d((1, 1)) + flatten (models/_resnet.py:149, 207-208): channels-last [B, C, H, W] -> [B, C]."""
    _lib.require_cuda(x)
    B, C, H, W = x.shape
    xc = x.float().contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_global_avgpool_nhwc(ptr(xc), B, H * W, C, ptr(y), stream_ptr(x.device)), "oess_global_avgpool_nhwc")
    return y


def planes_to_nhwc_padded(x, cp, stats=None):
    """[B, C, H, W] contiguous planes -> [B, cp, H, W] CHANNELS-LAST tensor whose channels C..cp-1 are zero; `stats`
    (float64 [3] from voxel.nonzero_standardize(phase=1)) applies the EventPreprocessor normalisation on the way."""
    _lib.require_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    y = torch.empty((B, cp, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        check(lib().oess_planes_to_nhwc_padded(ptr(x), B, C, H * W, ptr(stats), cp, ptr(y), stream_ptr(x.device)),
              "oess_planes_to_nhwc_padded")
    return y


def planes_to_nhwc_padded_w(x, cp, pad_w, stats=None):
    """planes_to_nhwc_padded with the rows zero-padded by pad_w pixels at both ends: returns the raw [B, H, W + 2 pad_w, cp]
    channels-last buffer (input of conv2d_rowunfold)."""
    _lib.require_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    y = torch.empty((B, H, W + 2 * pad_w, cp), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_planes_to_nhwc_padded_w(ptr(x), B, C, H, W, ptr(stats), cp, pad_w, ptr(y), stream_ptr(x.device)),
              "oess_planes_to_nhwc_padded_w")
    return y


def conv2d_pack_rowunfold(weight):
    """[Cout, Cin, KH, KW] -> [Cout, KH * roundup(KW * Cin, 32)], column (ky, kx * Cin + c): the K order of
    oess_conv2d_nhwc_tf32_rowunfold (thin-input convolution: the taps of a kernel row folded into the channel dimension)."""
    Cout, Cin, KH, KW = weight.shape
    cv = KW * Cin
    cvp = (cv + 31) // 32 * 32
    w = torch.zeros(Cout, KH, cvp, dtype=torch.float32, device=weight.device)
    w[:, :, :cv] = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, KH, cv)
    return round_tf32(w.reshape(Cout, KH * cvp).contiguous())


def conv2d_rowunfold(x_padded, w_packed, bias, KH, KW, W, relu=False, round_out=False):
    """Stride-1 'same' convolution of a thin channels-last input on the tensor cores.  x_padded: [B, H, W + KW - 1, Cin] from
    planes_to_nhwc_padded_w; returns [B, Cout, H, W] channels-last."""
    _lib.require_cuda(x_padded, w_packed, bias)
    B, H, Wp, Cin = x_padded.shape
    if Wp != W + KW - 1 or not x_padded.is_contiguous():
        raise ValueError("conv2d_rowunfold: x_padded must be a contiguous [B, H, W + KW - 1, Cin] tensor")
    Cout = w_packed.shape[0]
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x_padded.device, memory_format=torch.channels_last)
    bc = None if bias is None else _f32c(bias)
    with torch.cuda.device(x_padded.device):
        check(lib().oess_conv2d_nhwc_tf32_rowunfold(ptr(x_padded), ptr(w_packed), ptr(bc), ptr(y), B, H, W, Cin, Cout, KH, KW,
                                                    (1 if relu else 0) | (2 if round_out else 0), stream_ptr(x_padded.device)),
              "oess_conv2d_nhwc_tf32_rowunfold")
    return y


def conv_in(x, w_packed, bias, kernel_size, stride=1, padding=0, dilation=1, eps=1e-5, residual=None, relu=False):
    """act(InstanceNorm2d(conv(x) + bias) + residual) on the tensor cores, forward only (no autograd): per-sample
    statistics are accumulated in the conv's TMEM epilogue, one in-place pass normalises."""
    _lib.require_cuda(x, w_packed)
    B, Cin, H, W = x.shape
    KH = KW = int(kernel_size)
    Cout = w_packed.shape[0]
    cl = torch.channels_last
    xc = x.float().contiguous(memory_format=cl)
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=cl)
    res = None if residual is None else residual.float().contiguous(memory_format=cl)
    bc = None if bias is None else _f32c(bias)
    with torch.cuda.device(x.device):
        ws = _lib.workspace(8 * 2 * Cout * B, x.device)
        st = stream_ptr(x.device)
        check(lib().oess_conv2d_nhwc_tf32_instats(ptr(xc), ptr(w_packed), ptr(bc), ptr(y), B, H, W, Cin, Cout, KH, KW, stride,
                                                  padding, dilation, ptr(ws), st), "oess_conv2d_nhwc_tf32_instats")
        check(lib().oess_instancenorm_nhwc_sums(ptr(y), B, Ho * Wo, Cout, ptr(ws), float(eps), ptr(res), 1 if relu else 0, st),
              "oess_instancenorm_nhwc_sums")
    return y


def conv2d_dgrad_pack(weight, padding, dilation=1):
    """Input-gradient of a stride-1 convolution as a forward convolution: dX = conv(dY, W^T rotated by 180 degrees) with
    padding dilation * (K - 1) - padding.  Returns (packed weights for conv2d_tc, the padding to use)."""
    Cout, Cin, KH, KW = weight.shape
    if KH != KW:
        raise ValueError("square kernels only")
    wt = weight.detach().float().flip(2, 3).permute(1, 0, 2, 3).contiguous()      # [Cin, Cout, KH, KW]
    return conv2d_pack(wt), dilation * (KH - 1) - padding


def conv2d_wgrad(x, dy, kernel_size, padding=0, dilation=1):
    """dL/dW [Cout, Cin, K, K] of a stride-1 convolution on the tensor cores (split-K GEMM over pixels).
    x: [B, Cin, H, W], dy: [B, Cout, Ho, Wo]; both are used channels-last (copied if they are not)."""
    _lib.require_cuda(x, dy)
    xc = x.float().contiguous(memory_format=torch.channels_last)
    dyc = dy.float().contiguous(memory_format=torch.channels_last)
    B, Cin, H, W = xc.shape
    Cout = dyc.shape[1]
    K = int(kernel_size)
    dW = torch.empty((Cout, Cin, K, K), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_conv2d_wgrad_nhwc_tf32(ptr(xc), ptr(dyc), ptr(dW), B, H, W, Cin, Cout, K, K, int(padding),
                                                int(dilation), stream_ptr(x.device)), "oess_conv2d_wgrad_nhwc_tf32")
    return dW


class _ConvTC(torch.autograd.Function):
    """Stride-1 convolution whose forward, backward-data and backward-weight passes all run on the tcgen05 kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, padding, dilation):
        K = weight.shape[2]
        y = conv2d_tc(x, conv2d_pack(weight), bias, K, 1, padding, dilation)
        ctx.save_for_backward(x, weight)
        ctx.meta = (K, padding, dilation, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        K, padding, dilation, has_bias = ctx.meta
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            wp, pad = conv2d_dgrad_pack(weight, padding, dilation)
            dx = conv2d_tc(dy, wp, None, K, 1, pad, dilation)
        if ctx.needs_input_grad[1]:
            dW = conv2d_wgrad(x, dy, K, padding, dilation)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 2, 3))
        return dx, dW, db, None, None


def conv2d_tc_autograd(x, weight, bias=None, padding=0, dilation=1):
    """Differentiable stride-1 convolution on the tensor cores (TF32 operands, fp32 accumulate)."""
    if weight.shape[2] != weight.shape[3]:
        raise ValueError("square kernels only")
    return _ConvTC.apply(x, weight, bias, int(padding), int(dilation))


class _ConvINAct(torch.autograd.Function):
    """y = act(InstanceNorm2d(conv(x) + bias) + residual), stride 1, every pass on hand-written kernels: tcgen05 conv with
    per-sample statistics in its epilogue, in-place normalisation, and in the backward pass the InstanceNorm Jacobian
    (two HBM-bound kernels) followed by the tcgen05 backward-data and backward-weight convolutions."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, padding, dilation, eps, relu):
        B, Cin, H, W = x.shape
        Cout, _, K, _ = weight.shape
        cl = torch.channels_last
        xc = x.float().contiguous(memory_format=cl)
        Ho = H + 2 * padding - dilation * (K - 1)
        Wo = W + 2 * padding - dilation * (K - 1)
        xhat = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=cl)
        y = torch.empty_like(xhat)
        sums = torch.empty(B * 2 * Cout, dtype=torch.float64, device=x.device)
        res = None if residual is None else residual.float().contiguous(memory_format=cl)
        bc = None if bias is None else _f32c(bias)
        wp = conv2d_pack(weight)
        with torch.cuda.device(x.device):
            st = stream_ptr(x.device)
            check(lib().oess_conv2d_nhwc_tf32_instats(ptr(xc), ptr(wp), ptr(bc), ptr(xhat), B, H, W, Cin, Cout, K, K, 1,
                                                      padding, dilation, ptr(sums), st), "oess_conv2d_nhwc_tf32_instats")
            check(lib().oess_instancenorm_nhwc_sums_train(ptr(xhat), B, Ho * Wo, Cout, ptr(sums), float(eps), ptr(res),
                                                          1 if relu else 0, ptr(y), st), "oess_instancenorm_nhwc_sums_train")
        ctx.save_for_backward(xc, weight, xhat, y if relu else None, sums)
        ctx.meta = (K, padding, dilation, float(eps), bias is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, weight, xhat, y, sums = ctx.saved_tensors
        K, padding, dilation, eps, has_bias, has_res = ctx.meta
        B, Cout, Ho, Wo = xhat.shape
        cl = torch.channels_last
        dyc = dy.float().contiguous(memory_format=cl)
        dz = torch.empty_like(xhat)
        d_res = torch.empty_like(xhat) if (has_res and ctx.needs_input_grad[3]) else None
        bsums = torch.empty(B * 2 * Cout, dtype=torch.float64, device=dy.device)
        with torch.cuda.device(dy.device):
            check(lib().oess_instancenorm_nhwc_bwd(ptr(dyc), ptr(y), ptr(xhat), B, Ho * Wo, Cout, ptr(sums), ptr(bsums), eps,
                                                   ptr(dz), ptr(d_res), stream_ptr(dy.device)), "oess_instancenorm_nhwc_bwd")
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            wpd, pad = conv2d_dgrad_pack(weight, padding, dilation)
            dx = conv2d_tc(dz, wpd, None, K, 1, pad, dilation)
        if ctx.needs_input_grad[1]:
            dW = conv2d_wgrad(xc, dz, K, padding, dilation)
        if has_bias and ctx.needs_input_grad[2]:
            db = dz.sum(dim=(0, 2, 3))                       # exactly zero in exact arithmetic (a bias in front of InstanceNorm)
        return dx, dW, db, d_res, None, None, None, None


def conv_in_autograd(x, weight, bias=None, residual=None, padding=0, dilation=1, eps=1e-5, relu=False):
    """Differentiable act(InstanceNorm2d(conv(x) + bias) + residual) on hand-written kernels (see _ConvINAct)."""
    return _ConvINAct.apply(x, weight, bias, residual, int(padding), int(dilation), float(eps), bool(relu))


class _ConvBNAct(torch.autograd.Function):
    """y = act(BatchNorm2d_train(conv(x)) + residual), stride 1, bias-free conv: tcgen05 conv with batch statistics in its
    epilogue, fused normalise kernel (running statistics updated), and in the backward pass the BatchNorm Jacobian kernels
    + tcgen05 backward-data / backward-weight convolutions.  `bn` is the nn.BatchNorm2d module (buffers updated in place)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, bn, padding, dilation, relu):
        B, Cin, H, W = x.shape
        Cout, _, K, _ = weight.shape
        cl = torch.channels_last
        xc = x.float().contiguous(memory_format=cl)
        Ho = H + 2 * padding - dilation * (K - 1)
        Wo = W + 2 * padding - dilation * (K - 1)
        z = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=cl)
        y = torch.empty_like(z)
        res = None if residual is None else residual.float().contiguous(memory_format=cl)
        mom = _bn_momentum(bn)
        track = bn.track_running_stats and bn.running_mean is not None
        nb = ctypes.c_size_t(0)
        check(lib().oess_bn_ws_bytes(Cout, ctypes.byref(nb)), "oess_bn_ws_bytes")
        wp = conv2d_pack(weight)
        with torch.cuda.device(x.device):
            ws = _lib.workspace(nb.value, x.device)
            st = stream_ptr(x.device)
            check(lib().oess_conv2d_nhwc_tf32_stats(ptr(xc), ptr(wp), None, ptr(z), B, H, W, Cin, Cout, K, K, 1, padding,
                                                    dilation, ptr(ws), st), "oess_conv2d_nhwc_tf32_stats")
            sums = ws[:16 * Cout].view(torch.float64).clone()
            check(lib().oess_batchnorm_nhwc_sums_train(ptr(z), B * Ho * Wo, Cout, ptr(gamma), ptr(beta),
                                                       ptr(bn.running_mean) if track else None,
                                                       ptr(bn.running_var) if track else None, float(bn.eps), mom, ptr(res),
                                                       1 if relu else 0, ptr(y), ptr(ws), ws.numel(), st),
                  "oess_batchnorm_nhwc_sums_train")
            if track and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        ctx.save_for_backward(xc, weight, z, y if relu else None, sums, gamma)
        ctx.meta = (K, padding, dilation, float(bn.eps), residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, weight, z, y, sums, gamma = ctx.saved_tensors
        K, padding, dilation, eps, has_res = ctx.meta
        B, Cout, Ho, Wo = z.shape
        dyc = dy.float().contiguous(memory_format=torch.channels_last)
        dz = torch.empty_like(z)
        d_res = torch.empty_like(z) if (has_res and ctx.needs_input_grad[4]) else None
        bsums = torch.empty(2 * Cout, dtype=torch.float64, device=dy.device)
        with torch.cuda.device(dy.device):
            check(lib().oess_batchnorm_nhwc_bwd(ptr(dyc), ptr(y), ptr(z), B * Ho * Wo, Cout, ptr(sums), ptr(gamma), ptr(bsums),
                                                eps, ptr(dz), ptr(d_res), stream_ptr(dy.device)), "oess_batchnorm_nhwc_bwd")
        dx = dW = None
        if ctx.needs_input_grad[0]:
            wpd, pad = conv2d_dgrad_pack(weight, padding, dilation)
            dx = conv2d_tc(dz, wpd, None, K, 1, pad, dilation)
        if ctx.needs_input_grad[1]:
            dW = conv2d_wgrad(xc, dz, K, padding, dilation)
        dgamma = bsums[Cout:].float() if ctx.needs_input_grad[2] else None
        dbeta = bsums[:Cout].float() if ctx.needs_input_grad[3] else None
        return dx, dW, dgamma, dbeta, d_res, None, None, None, None


def conv_bn_autograd(x, conv, bn, residual=None, relu=False):
    """Differentiable act(bn(conv(x)) + residual) for a bias-free stride-1 `conv` and a train-mode nn.BatchNorm2d `bn`."""
    if conv.bias is not None or conv.stride != (1, 1):
        raise ValueError("conv_bn_autograd: bias-free stride-1 convolutions only")
    return _ConvBNAct.apply(x, conv.weight, bn.weight, bn.bias, residual, bn, conv.padding[0], conv.dilation[0], bool(relu))


def unsharp_rescale(img, kernel5x5, amount, imin, imax, quantize=True):
    """PostProcessor of the reconstruction (e2vid/image_reconstructor.py:126-140): unsharp mask + intensity rescaling +
    8-bit quantisation in one kernel.  img: [B, 1, H, W] float32 CUDA; kernel5x5: 25 float32 taps on the device."""
    _lib.require_cuda(img, kernel5x5)
    if img.ndim != 4 or img.shape[1] != 1 or img.dtype != torch.float32:
        raise ValueError("unsharp_rescale: img must be float32 [B, 1, H, W]")
    img = img.contiguous()
    k = kernel5x5.to(img.device, torch.float32).contiguous()
    if k.numel() != 25:
        raise ValueError("unsharp_rescale: the Gaussian kernel must have 5 x 5 taps")
    B, _, H, W = img.shape
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().oess_unsharp_rescale(_lib.ptr(img), _lib.ptr(k), B, H, W, float(amount), float(imin), float(imax),
                                                   int(bool(quantize)), _lib.ptr(out), _lib.stream_ptr(img.device)),
                   "oess_unsharp_rescale")
    return out
