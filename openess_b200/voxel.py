"""Batched, device-level voxelisation API (torch CUDA tensors in, torch CUDA tensors out).

Thin host logic over the C ABI: argument checking, workspace, stream plumbing.  One call voxelises F
event-frames; `frame_offsets` (int64 [F+1]) delimits the frames inside the concatenated event arrays.
"""
import os

import torch

from . import _lib
from ._lib import KIND_TBILINEAR, KIND_TRILINEAR, MODE_ATOMIC, MODE_ORDERED, check, lib, ptr, require_cuda, stream_ptr

_MODES = {"ordered": MODE_ORDERED, "atomic": MODE_ATOMIC, MODE_ORDERED: MODE_ORDERED, MODE_ATOMIC: MODE_ATOMIC}


def resolve_mode(mode=None):
    """'ordered' (bit-exact with the reference, default) or 'atomic' (fast, ~1e-6 tolerance)."""
    if mode is None:
        mode = os.environ.get("OPENESS_B200_VOXEL_MODE", "ordered")
    try:
        return _MODES[mode]
    except KeyError:
        raise ValueError(f"unknown voxel mode {mode!r}; use 'ordered' or 'atomic'")


def _offsets(frame_offsets, n, device):
    if frame_offsets is None:
        return torch.tensor([0, n], dtype=torch.int64, device=device)
    if not torch.is_tensor(frame_offsets):
        frame_offsets = torch.as_tensor(frame_offsets, dtype=torch.int64)
    fo = frame_offsets.to(device=device, dtype=torch.int64).contiguous()
    if fo.ndim != 1 or fo.numel() < 1:
        raise ValueError("frame_offsets must be 1-D with F+1 entries")
    return fo


def voxel_trilinear(x, y, pol, t, C, H, W, frame_offsets=None, mode=None, normalize=False, out=None):
    """VoxelGrid.convert semantics (DSEC/dataset/representations.py:15-55) for F frames -> [F, C, H, W]."""
    require_cuda(x, y, pol, t)
    dev = x.device
    n = x.numel()
    for a in (x, y, pol, t):
        if a.dtype != torch.float32 or a.ndim != 1 or a.numel() != n or not a.is_contiguous():
            raise ValueError("x, y, pol, t must be contiguous 1-D float32 tensors of equal length")
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    m = resolve_mode(mode)
    if out is None:
        out = torch.empty((F, C, H, W), dtype=torch.float32, device=dev)
    elif out.shape != (F, C, H, W) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise ValueError("out must be a contiguous float32 [F, C, H, W] tensor on the input device")
    with torch.cuda.device(dev):
        nbytes = _lib.voxel_ws_bytes(KIND_TRILINEAR, m, n, F, C, H, W)
        ws = _lib.workspace(nbytes, dev)
        check(lib().oess_voxel_trilinear(ptr(x), ptr(y), ptr(pol), ptr(t), ptr(fo), n, F, C, H, W, m,
                                         int(bool(normalize)), ptr(out), ptr(ws), ws.numel(), stream_ptr(dev)),
              "oess_voxel_trilinear")
    return out


def _ev4_fn(ev4, base):
    if ev4.dtype == torch.int64:
        return getattr(lib(), base + "_i64")
    if ev4.dtype == torch.float64:
        return getattr(lib(), base + "_f64")
    raise TypeError(f"events must be int64 or float64 [N, 4], got {ev4.dtype}")


def voxel_tbilinear(ev4, C, H, W, frame_offsets=None, separate_pol=True, mode=None, mutate_p=True, out=None):
    """generate_voxel_grid semantics (datasets/data_util.py:51-117) for F frames -> [F, C or 2C, H, W]."""
    require_cuda(ev4)
    if ev4.ndim != 2 or ev4.shape[1] != 4 or not ev4.is_contiguous():
        raise ValueError("events must be a contiguous [N, 4] tensor (x, y, t, p)")
    dev = ev4.device
    n = ev4.shape[0]
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    m = resolve_mode(mode)
    planes = 2 * C if separate_pol else C
    if out is None:
        out = torch.empty((F, planes, H, W), dtype=torch.float32, device=dev)
    fn = _ev4_fn(ev4, "oess_voxel_tbilinear")
    with torch.cuda.device(dev):
        nbytes = _lib.voxel_ws_bytes(KIND_TBILINEAR, m, n, F, C, H, W)
        ws = _lib.workspace(nbytes, dev)
        check(fn(ptr(ev4), ptr(fo), n, F, C, H, W, int(bool(separate_pol)), m, int(bool(mutate_p)), ptr(out),
                 ptr(ws), ws.numel(), stream_ptr(dev)), "oess_voxel_tbilinear")
    return out


def voxel_histogram(ev4, H, W, frame_offsets=None, mutate_p=True, out=None, status=None):
    """generate_event_histogram semantics (datasets/data_util.py:17-35) -> [F, 2, H, W] (neg, pos)."""
    require_cuda(ev4)
    if ev4.ndim != 2 or ev4.shape[1] != 4 or not ev4.is_contiguous():
        raise ValueError("events must be a contiguous [N, 4] tensor (x, y, t, p)")
    dev = ev4.device
    n = ev4.shape[0]
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    if out is None:
        out = torch.empty((F, 2, H, W), dtype=torch.float32, device=dev)
    fn = _ev4_fn(ev4, "oess_voxel_histogram")
    with torch.cuda.device(dev):
        check(fn(ptr(ev4), ptr(fo), n, F, H, W, int(bool(mutate_p)), ptr(out), ptr(status), stream_ptr(dev)),
              "oess_voxel_histogram")
    return out


def _ddd17_check(t, xyp):
    require_cuda(t, xyp)
    if t.dtype != torch.int64 or xyp.dtype != torch.int16:
        raise TypeError("DDD17 records are t int64 [N] / [N, 1] and xyp int16 [N, 3] (events.dat.t / events.dat.xyp)")
    n = xyp.shape[0]
    if xyp.ndim != 2 or xyp.shape[1] != 3 or t.numel() != n or not t.is_contiguous() or not xyp.is_contiguous():
        raise ValueError("DDD17 records: contiguous t with N elements and contiguous xyp [N, 3]")
    return n


def voxel_tbilinear_ddd17(t, xyp, C, H, W, frame_offsets=None, separate_pol=True, mode=None, out=None):
    """generate_voxel_grid (datasets/data_util.py:51-117) straight from the DDD17 on-disk records (14 B / event): bit-equal
    to voxel_tbilinear on the int64 rows example_loader_ddd17.py:41-54 assembles from them."""
    n = _ddd17_check(t, xyp)
    dev = xyp.device
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    m = resolve_mode(mode)
    if out is None:
        out = torch.empty((F, 2 * C if separate_pol else C, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _lib.workspace(_lib.voxel_ws_bytes(KIND_TBILINEAR, m, n, F, C, H, W), dev)
        check(lib().oess_voxel_tbilinear_ddd17(ptr(t), ptr(xyp), ptr(fo), n, F, C, H, W, int(bool(separate_pol)), m, ptr(out),
                                               ptr(ws), ws.numel(), stream_ptr(dev)), "oess_voxel_tbilinear_ddd17")
    return out


def voxel_histogram_ddd17(t, xyp, H, W, frame_offsets=None, out=None, status=None):
    """generate_event_histogram (datasets/data_util.py:17-35) from the DDD17 on-disk records -> [F, 2, H, W] (neg, pos)."""
    n = _ddd17_check(t, xyp)
    dev = xyp.device
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    if out is None:
        out = torch.empty((F, 2, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().oess_voxel_histogram_ddd17(ptr(t), ptr(xyp), ptr(fo), n, F, H, W, ptr(out), ptr(status), stream_ptr(dev)),
              "oess_voxel_histogram_ddd17")
    return out


def dsec_rectify_tnorm(x, y, t, p, rectify_map, frame_offsets=None, status=None, out=None):
    """sequence_ov.py:204-210 + :154-159 for F frames: raw (u16 x, u16 y, i64|u32 t, u8 p) -> f32 x', y', pol, t.

    t may be int64 microseconds (t_offset added, what EventSlicer returns) or the on-disk uint32 timestamps."""
    require_cuda(x, y, t, p, rectify_map)
    dev = x.device
    n = x.numel()
    if x.dtype != torch.uint16 or y.dtype != torch.uint16 or p.dtype != torch.uint8 or \
            t.dtype not in (torch.int64, torch.uint32):
        raise TypeError("raw DSEC records are x, y uint16; t int64 or uint32 (microseconds); p uint8")
    if rectify_map.dtype != torch.float32 or rectify_map.ndim != 3 or rectify_map.shape[2] != 2:
        raise ValueError("rectify_map must be float32 [H, W, 2]")
    H, W = rectify_map.shape[:2]
    fo = _offsets(frame_offsets, n, dev)
    F = fo.numel() - 1
    if out is None:
        out = tuple(torch.empty(n, dtype=torch.float32, device=dev) for _ in range(4))
    xo, yo, po, to = out
    fn = lib().oess_dsec_rectify_tnorm if t.dtype == torch.int64 else lib().oess_dsec_rectify_tnorm_u32
    with torch.cuda.device(dev):
        check(fn(ptr(x.contiguous()), ptr(y.contiguous()), ptr(t.contiguous()), ptr(p.contiguous()),
                 ptr(rectify_map.contiguous()), ptr(fo), n, F, H, W, ptr(xo), ptr(yo), ptr(po), ptr(to), ptr(status),
                 stream_ptr(dev)), "oess_dsec_rectify_tnorm")
    return xo, yo, po, to


def dsec_events_to_voxel_grid(x, y, t, p, rectify_map, C, frame_offsets=None, mode=None, normalize=False, out=None,
                              scratch=None, status=None):
    """Raw DSEC records of F frames -> [F, C, H, W] voxel grids on the device: rectify_events +
    events_to_voxel_grid + VoxelGrid.convert (sequence_ov.py:204-223, representations.py:15-55) in two ABI calls.
    This is the GPU-side sample assembly of SURVEY.md 7.1 step 3 (F = B * nr_events_data frames per call)."""
    H, W = rectify_map.shape[:2]
    xo, yo, po, to = dsec_rectify_tnorm(x, y, t, p, rectify_map, frame_offsets, status=status, out=scratch)
    return voxel_trilinear(xo, yo, po, to, C, H, W, frame_offsets=frame_offsets, mode=mode, normalize=normalize, out=out)


def nonzero_standardize(x, n_groups=1, unbiased=False, phase=0, stats=None):
    """In-place nonzero mean/std standardisation of a float32 CUDA tensor viewed as [n_groups, -1]."""
    require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be a contiguous float32 tensor")
    if x.numel() == 0:
        return x, stats
    if x.numel() % n_groups:
        raise ValueError("numel not divisible by n_groups")
    if stats is None:
        stats = torch.empty((n_groups, 3), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().oess_nonzero_standardize(ptr(x), x.numel() // n_groups, n_groups, ptr(stats), phase,
                                             int(bool(unbiased)), stream_ptr(x.device)), "oess_nonzero_standardize")
    return x, stats
