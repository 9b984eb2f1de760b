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


def allreduce_confusion_(conf):
    """Integer confusion matrices add exactly across ranks (validation, metrics.py)."""
    return allreduce_sum_(conf)
