"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The hot path shards over independent units (event-frames / samples), so voxelisation needs no
data-path collective; the only collectives are the per-step gradient all-reduce of the trainable modules and
the optional tiny all-reduces that give exact global-batch semantics for batch-coupled statistics
(EventPreprocessor sums, Dice/CE partial sums)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, local, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local) if device is None else device
        dist.init_process_group(backend, **kw)
    return rank, local, world


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(n_units, rank, world):
    """Contiguous, balanced [lo, hi) of `n_units` independent units for `rank` (first n % world ranks get +1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, rem = divmod(int(n_units), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(frame_offsets, rank, world):
    """Slice a concatenated event batch: returns (frame_lo, frame_hi, event_lo, event_hi, local_offsets)."""
    F = len(frame_offsets) - 1
    lo, hi = shard_range(F, rank, world)
    ev_lo, ev_hi = int(frame_offsets[lo]), int(frame_offsets[hi])
    local = [int(o) - ev_lo for o in frame_offsets[lo:hi + 1]]
    return lo, hi, ev_lo, ev_hi, local


def allreduce_sum_(t):
    """In-place sum over ranks (identity for a single process).  Used for float64 partial sums / int64 counts,
    where the reduction is exact or order-insensitive enough to give global-batch semantics."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_gradients(params, bucket_bytes=32 << 20, average=True):
    """Bucketed gradient all-reduce of the trainable parameters (flatten -> all_reduce -> unflatten).

    Parameters whose .grad is None are skipped: SemSegE2VID.decoder_scale_5 and DeepLabHead.pixel_feature
    never receive gradients in the reference's forward, which would stall a hook-based DDP.
    Buckets follow registration order, so every rank builds identical buckets."""
    w = world_size()
    if w == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    n_calls, i = 0, 0
    while i < len(grads):
        j, size = i, 0
        while j < len(grads) and (j == i or size + grads[j].numel() * grads[j].element_size() <= bucket_bytes) \
                and grads[j].dtype == grads[i].dtype:
            size += grads[j].numel() * grads[j].element_size()
            j += 1
        flat = torch.cat([g.reshape(-1) for g in grads[i:j]])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(w)
        off = 0
        for g in grads[i:j]:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_calls += 1
        i = j
    return n_calls


class GradientReducer:
    """Bucketed gradient all-reduce OVERLAPPED with backward ("bucketed and overlapped with backward").

    Parameters are cut into buckets of `bucket_bytes` in REVERSE registration order (the order backward produces
    gradients in); a post-accumulate-grad hook copies each finished gradient into its bucket's flat buffer, and a bucket
    whose gradients are all in is all-reduced asynchronously (buckets are always launched in index order, so every rank
    issues the same collective sequence) while backward keeps running.  `finish()` waits, averages and scatters the reduced
    values back into the `.grad` tensors.  Parameters that never receive a gradient (SemSegE2VID.decoder_scale_5,
    DeepLabHead.pixel_feature: unused in the reference's forward) are detected on the first step, which runs the plain
    post-backward `allreduce_gradients`, and are left out of the buckets -- their `.grad` stays None, so AdamW skips them
    exactly as in the single-GPU reference."""

    def __init__(self, params, bucket_bytes=8 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes, self.average = int(bucket_bytes), average
        self.buckets = None           # list of dicts: params, flat, offsets, pending, launched, work
        self._hooks = []
        self._index = {}
        self.stats = {"buckets": 0, "bytes": 0, "overlapped_calls": 0}

    def _build(self):
        live = [p for p in self.params if p.grad is not None]
        self.buckets, cur, size = [], [], 0
        for p in reversed(live):
            nb = p.numel() * p.element_size()
            if cur and (size + nb > self.bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nb
        if cur:
            self.buckets.append(cur)
        built = []
        for bi, ps in enumerate(self.buckets):
            n = sum(p.numel() for p in ps)
            flat = torch.empty(n, dtype=ps[0].dtype, device=ps[0].device)
            offs, o = [], 0
            for p in ps:
                offs.append(o)
                o += p.numel()
            built.append({"params": ps, "flat": flat, "offsets": offs, "pending": len(ps), "work": None})
            for j, p in enumerate(ps):
                self._index[p] = (bi, j)
        self.buckets = built
        self.stats["buckets"] = len(built)
        self.stats["bytes"] = sum(b["flat"].numel() * b["flat"].element_size() for b in built)
        for p in live:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p):
        bi, j = self._index[p]
        b = self.buckets[bi]
        o = b["offsets"][j]
        b["flat"][o:o + p.numel()].copy_(p.grad.reshape(-1))
        b["pending"] -= 1
        self._launch_ready()

    def _launch_ready(self):
        while self._next < len(self.buckets) and self.buckets[self._next]["pending"] <= 0:
            b = self.buckets[self._next]
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, async_op=True)
            self.stats["overlapped_calls"] += 1
            self._next += 1

    def prepare(self):
        """Call before backward()."""
        if self.buckets is None:
            return
        self._next = 0
        for b in self.buckets:
            b["pending"], b["work"] = len(b["params"]), None

    def finish(self):
        """Call after backward(): returns the number of all-reduce calls of this step."""
        w = world_size()
        if w == 1:
            return 0
        if self.buckets is None:                 # first step: learn which parameters receive gradients
            n = allreduce_gradients(self.params, self.bucket_bytes, self.average)
            self._build()
            return n
        for b in self.buckets:                   # a gradient that did not arrive this step counts as zero
            if b["pending"] > 0:
                for p, o in zip(b["params"], b["offsets"]):
                    if p.grad is None:
                        b["flat"][o:o + p.numel()].zero_()
                b["pending"] = 0
        self._launch_ready()
        for b in self.buckets:
            b["work"].wait()
            if self.average:
                b["flat"].div_(w)
            grads = [p.grad for p in b["params"] if p.grad is not None]
            views = [b["flat"][o:o + p.numel()].view_as(p) for p, o in zip(b["params"], b["offsets"]) if p.grad is not None]
            if grads:
                torch._foreach_copy_(grads, views)
        return len(self.buckets)


def allreduce_confusion_(conf):
    """Integer confusion matrices add exactly across ranks (validation, metrics.py)."""
    return allreduce_sum_(conf)
