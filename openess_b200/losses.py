"""Device-level loss-side ops over the C ABI (torch CUDA tensors in/out) + autograd wrappers.

  segpool     training/pretrain_trainer.py:445-465   superpixel mean-pool (replaces the sparse one-hot matmul)
  infonce     utils/loss_functions.py:138-153        NCELoss
  dice_ce     utils/loss_functions.py:6-24,96-135    TaskLoss = DiceLoss + CrossEntropyLoss(ignore_index)
  confusion   evaluation/metrics.py:4-23             semseg_compute_confusion
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------ segpool
def segpool_forward(feat, seg, S, M, status=None):
    """feat [B,Cf,H,W] f32, seg [B,H,W] int64 -> (pooled [M,Cf], counts [M])."""
    require_cuda(feat, seg)
    feat = _f32c(feat)
    seg = seg.to(torch.int64).contiguous()
    B, Cf, H, W = feat.shape
    pooled = torch.empty((M, Cf), dtype=torch.float32, device=feat.device)
    counts = torch.empty((M,), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(lib().oess_segpool_fwd(ptr(feat), ptr(seg), B, Cf, H, W, int(S), int(M), ptr(pooled), ptr(counts),
                                     ptr(status), None, 0, stream_ptr(feat.device)), "oess_segpool_fwd")
    return pooled, counts


def segpool_backward(d_pooled, seg, counts, S, shape):
    B, Cf, H, W = shape
    d_pooled = _f32c(d_pooled)
    d_feat = torch.empty(shape, dtype=torch.float32, device=d_pooled.device)
    with torch.cuda.device(d_pooled.device):
        check(lib().oess_segpool_bwd(ptr(d_pooled), ptr(seg), ptr(counts), B, Cf, H, W, int(S), int(d_pooled.shape[0]),
                                     ptr(d_feat), stream_ptr(d_pooled.device)), "oess_segpool_bwd")
    return d_feat


class _SegPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, seg, S, M):
        seg = seg.to(torch.int64).contiguous()
        pooled, counts = segpool_forward(feat, seg, S, M)
        ctx.save_for_backward(seg, counts)
        ctx.S, ctx.shape = S, tuple(feat.shape)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        seg, counts = ctx.saved_tensors
        return segpool_backward(d_pooled, seg, counts, ctx.S, ctx.shape), None, None, None


def superpixel_pool(feat, superpixels, superpixel_size, M=None):
    """pretrain_trainer.py:445-465: ids += b*S; pooled[m] = sum_{pix: id == m} feat[pix] / (count[m] + 1e-6).

    M defaults to max(id') + 1, which is what torch.sparse_coo_tensor infers in the reference (one host sync,
    like the reference's sparse constructor)."""
    B = feat.shape[0]
    if M is None:
        off = torch.arange(0, B * superpixel_size, superpixel_size, device=superpixels.device)[:, None, None]
        M = int((superpixels + off).max().item()) + 1
    return _SegPool.apply(feat, superpixels, int(superpixel_size), int(M))


# ------------------------------------------------------------------------------------------ InfoNCE
def _infonce_raw(k, q, temperature, need_grad):
    require_cuda(k, q)
    k, q = _f32c(k), _f32c(q)
    M, D = k.shape
    nbytes = ctypes.c_size_t(0)
    check(lib().oess_infonce_ws_bytes(M, D, ctypes.byref(nbytes)), "oess_infonce_ws_bytes")
    loss = torch.empty(1, dtype=torch.float32, device=k.device)
    dk = torch.empty_like(k) if need_grad else None
    dq = torch.empty_like(q) if need_grad else None
    with torch.cuda.device(k.device):
        ws = _lib.workspace(nbytes.value, k.device)
        check(lib().oess_infonce(ptr(k), ptr(q), M, D, float(temperature), ptr(loss), ptr(dk), ptr(dq), ptr(ws),
                                 ws.numel(), stream_ptr(k.device)), "oess_infonce")
    return loss, dk, dq


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, k, q, temperature):
        need = k.requires_grad or q.requires_grad
        loss, dk, dq = _infonce_raw(k, q, temperature, need)
        if need:
            ctx.save_for_backward(dk, dq)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        dk, dq = ctx.saved_tensors
        return dk * g, dq * g, None


def infonce(k, q, temperature):
    """loss_functions.py:147-153: CrossEntropy((k @ q.T) / T, arange(M)), mean reduction."""
    if k.shape != q.shape or k.ndim != 2:
        raise ValueError("k and q must both be [M, D]")
    return _InfoNCE.apply(k, q, float(temperature))


# ------------------------------------------------------------------------------------------ Dice + CE
def dice_ce_partials(logits, target, ignore_index):
    """One pass over the logits -> float64 [2K+3] = {inter[K], denom[K], ce_sum, n_valid, n_bad} (all-reducible); n_bad counts
    targets that are neither a class nor `ignore_index`."""
    require_cuda(logits, target)
    logits = _f32c(logits)
    target = target.to(torch.int64).contiguous()
    B, K, H, W = logits.shape
    partials = torch.empty(2 * K + 3, dtype=torch.float64, device=logits.device)
    ig = -(1 << 62) if ignore_index is None else int(ignore_index)
    with torch.cuda.device(logits.device):
        check(lib().oess_dice_ce_partials(ptr(logits), ptr(target), B, K, H, W, ig, ptr(partials),
                                          stream_ptr(logits.device)), "oess_dice_ce_partials")
    return partials


def dice_ce_finish(partials, K, w_dice, w_ce, ignore_index=None):
    losses = torch.empty(3, dtype=torch.float32, device=partials.device)
    ig = -(1 << 62) if ignore_index is None else int(ignore_index)
    with torch.cuda.device(partials.device):
        check(lib().oess_dice_ce_finish_ex(ptr(partials), K, ig, float(w_dice), float(w_ce), ptr(losses),
                                           stream_ptr(partials.device)), "oess_dice_ce_finish_ex")
    return losses


class _LabelCheck:
    """Out-of-range labels (neither a class nor ignore_index) make the reference raise (scatter_ / CrossEntropyLoss).  The fused
    kernel counts them; the count of call i is copied to pinned host memory without a synchronisation and inspected at call
    i + 1 (or by `flush()`), so corrupt labels stop the run one step late instead of training silently."""

    def __init__(self):
        self._pending = None            # (pinned host tensor, event)
        self._bufs, self._turn = None, 0

    def submit(self, partials, K):
        self.check(wait=False)
        if self._bufs is None:          # two pinned doubles, reused in turn (no cudaHostAlloc per step)
            self._bufs = [torch.empty(1, dtype=torch.float64).pin_memory() for _ in range(2)]
        if self._pending is not None:   # the previous count has not arrived yet: wait for it rather than overwrite its buffer
            self.check(wait=True)
        host = self._bufs[self._turn]
        self._turn ^= 1
        host.copy_(partials[2 * K + 2:2 * K + 3], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(partials.device))
        self._pending = (host, ev)

    def check(self, wait=True):
        if self._pending is None:
            return
        host, ev = self._pending
        if not wait and not ev.query():
            return                      # not there yet: looked at again on the next call
        ev.synchronize()
        self._pending = None
        if float(host[0]) != 0.0:
            raise ValueError(f"dice_ce: {int(host[0])} target value(s) outside [0, K) that are not ignore_index "
                             "(utils/loss_functions.py:43-57 make_one_hot / CrossEntropyLoss raise on these)")


label_check = _LabelCheck()


class _DiceCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index, w_dice, w_ce, reduce_partials):
        logits = _f32c(logits)
        target = target.to(torch.int64).contiguous()
        partials = dice_ce_partials(logits, target, ignore_index)
        if reduce_partials is not None:          # exact global-batch semantics across ranks
            reduce_partials(partials)
        K = logits.shape[1]
        losses = dice_ce_finish(partials, K, w_dice, w_ce, ignore_index)
        label_check.submit(partials, K)
        ctx.save_for_backward(logits, target, partials)
        ctx.cfg = (ignore_index, w_dice, w_ce)
        # With all-reduced partial sums every rank differentiates the GLOBAL loss w.r.t. its LOCAL logits, so the true
        # parameter gradient is the SUM over ranks; the data-parallel step AVERAGES gradients (parallel.GradientReducer /
        # allreduce_gradients, average=True), hence the factor world_size here (without it the Dice/CE term
        # is down-weighted by 1 / world_size against the locally computed contrastive loss).
        from . import parallel as _parallel
        ctx.grad_scale = float(_parallel.world_size()) if reduce_partials is not None else 1.0
        return losses[2]

    @staticmethod
    def backward(ctx, g):
        logits, target, partials = ctx.saved_tensors
        ignore_index, w_dice, w_ce = ctx.cfg
        B, K, H, W = logits.shape
        d = torch.empty_like(logits)
        gs = g.reshape(1).to(torch.float32).contiguous()
        if ctx.grad_scale != 1.0:
            gs = gs * ctx.grad_scale
        ig = -(1 << 62) if ignore_index is None else int(ignore_index)
        with torch.cuda.device(logits.device):
            check(lib().oess_dice_ce_bwd(ptr(logits), ptr(target), B, K, H, W, ig, ptr(partials), float(w_dice),
                                         float(w_ce), ptr(gs), ptr(d), stream_ptr(logits.device)), "oess_dice_ce_bwd")
        return d, None, None, None, None, None


def dice_ce(logits, target, ignore_index=255, w_dice=1.0, w_ce=1.0, reduce_partials=None):
    """w_dice * DiceLoss + w_ce * CrossEntropyLoss(ignore_index) in one fused pass (+ one fused backward)."""
    return _DiceCE.apply(logits, target, ignore_index, float(w_dice), float(w_ce), reduce_partials)


# ------------------------------------------------------------------------------------------ confusion
def confusion(pred, gt, num_classes, ignore_label, out=None, status=None):
    """metrics.py:4-23 -> int64 [K,K] conf[gt, pred]; accumulates into `out` if given."""
    require_cuda(pred, gt)
    pred = pred.to(torch.int64).contiguous()
    gt = gt.to(torch.int64).contiguous()
    if pred.numel() != gt.numel():
        raise ValueError("pred and gt must have the same number of elements")
    if out is None:
        out = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        check(lib().oess_confusion(ptr(pred), ptr(gt), pred.numel(), int(num_classes), int(ignore_label), ptr(out),
                                   ptr(status), stream_ptr(pred.device)), "oess_confusion")
    return out


def argmax_confusion(logits, gt, ignore_label, out=None, status=None):
    """base_trainer_ov.py:463-466 + metrics.py:4-23 fused: conf[gt, argmax_c logits] += 1 over gt != ignore, without
    materialising the prediction map.  logits [B, K, H, W] float32, gt int64 [B, H, W]; accumulates into `out`."""
    require_cuda(logits, gt)
    logits = _f32c(logits)
    gt = gt.to(torch.int64).contiguous()
    B, K, H, W = logits.shape
    if gt.numel() != B * H * W:
        raise ValueError("gt must have B*H*W elements")
    if out is None:
        out = torch.zeros((K, K), dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        check(lib().oess_argmax_confusion(ptr(logits), ptr(gt), B, K, H, W, int(ignore_label), ptr(out), ptr(status),
                                          stream_ptr(logits.device)), "oess_argmax_confusion")
    return out


# ------------------------------------------------------------------------------------------ consistency losses
class _L1Mean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32c(a), _f32c(b)
        require_cuda(a, b)
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        acc = torch.empty(1, dtype=torch.float64, device=a.device)
        with torch.cuda.device(a.device):
            check(lib().oess_l1_mean(ptr(a), ptr(b), a.numel(), ptr(loss), ptr(acc), stream_ptr(a.device)), "oess_l1_mean")
        ctx.save_for_backward(a, b)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        gs = g.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            check(lib().oess_l1_mean_bwd(ptr(a), ptr(b), a.numel(), ptr(gs), ptr(da), ptr(db), stream_ptr(a.device)),
                  "oess_l1_mean_bwd")
        return da, db


def l1_mean(a, b):
    """torch.nn.L1Loss()(a, b) (openess_trainer.py:456): mean |a - b|, fused forward and backward."""
    if a.shape != b.shape:
        raise ValueError("shapes differ")
    return _L1Mean.apply(a, b)


class _CosConsistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32c(a), _f32c(b)
        require_cuda(a, b)
        B, K = a.shape[:2]
        HW = a.numel() // (B * K)
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        acc = torch.empty(1, dtype=torch.float64, device=a.device)
        with torch.cuda.device(a.device):
            check(lib().oess_cos_consistency(ptr(a), ptr(b), B, K, HW, ptr(loss), ptr(acc), stream_ptr(a.device)),
                  "oess_cos_consistency")
        ctx.save_for_backward(a, b)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        B, K = a.shape[:2]
        HW = a.numel() // (B * K)
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        gs = g.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(a.device):
            check(lib().oess_cos_consistency_bwd(ptr(a), ptr(b), B, K, HW, ptr(gs), ptr(da), ptr(db),
                                                 stream_ptr(a.device)), "oess_cos_consistency_bwd")
        return da, db


def cosine_consistency(a, b):
    """mean(1 - cosine_similarity(a, b, dim=1)) for [B, K, H, W] maps (openess_trainer.py:460)."""
    if a.shape != b.shape or a.ndim < 2:
        raise ValueError("a and b must have the same [B, K, ...] shape")
    return _CosConsistency.apply(a, b)


# ------------------------------------------------------------------------------------------ ConvLSTM gates
def convlstm_gates(gates, prev_cell=None):
    """Fused pointwise tail of ConvLSTM.forward (e2vid/model/submodules.py:203-212) -> (hidden, cell). No grad:
    the E2VID encoder is frozen and run under no_grad in every OpenESS trainer."""
    require_cuda(gates)
    gates = _f32c(gates)
    B, C4 = gates.shape[:2]
    C = C4 // 4
    HW = gates.numel() // (B * C4)
    hidden = torch.empty((B, C) + tuple(gates.shape[2:]), dtype=torch.float32, device=gates.device)
    cell = torch.empty_like(hidden)
    pc = None if prev_cell is None else _f32c(prev_cell)
    with torch.cuda.device(gates.device):
        check(lib().oess_convlstm_gates(ptr(gates), ptr(pc), ptr(hidden), ptr(cell), B, C, HW, stream_ptr(gates.device)),
              "oess_convlstm_gates")
    return hidden, cell
