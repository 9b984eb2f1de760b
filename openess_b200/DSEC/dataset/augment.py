"""Training-time augmentation of a DSEC batch on the device: the `self.augmentation` branch of
`Sequence.__getitem__` (DSEC/dataset/sequence_ov.py:362-382 recon2voxel, :387-407 frame2voxel) for a whole batch.

The reference draws, per sample and in this order, with Python's `random`: flip (`random() >= 0.5`), then for each of
brightness / contrast / noise a gate (`random() >= 0.5`) followed -- only if the gate is open -- by `uniform(0.8, 1.2)`
(brightness, contrast) or `torch.randn(frame.size()) * 0.05` (noise).  `draw_params` replays exactly that sequence of
`random` calls per sample, so a loader seeded like the reference's worker takes the same decisions; the noise tensor
itself comes from the caller's generator on the device (same distribution, not the same stream as torch's CPU generator).

`augment_batch_` applies the decisions in place: `oess_hflip_rows` on the event tensor, frame (or reconstruction), label,
pseudo-label and superpixel maps, `oess_frame_color_aug` (brightness -> contrast -> noise, torchvision semantics) on the
frame.  The 256 x 64 x 64 `sam_feat` placeholder of the reference is all ones (:360): flipping it is the identity.
"""
import random as _random

import torch

from ... import _lib
from ..._lib import check, lib, ptr, stream_ptr


def draw_params(batch_size, rng=_random):
    """-> dict of host lists (flip: bool, brightness / contrast: float with 1.0 = off, noise: bool), one entry per sample, drawn
    with the reference's call sequence (sequence_ov.py:388-406)."""
    out = {"flip": [], "brightness": [], "contrast": [], "noise": []}
    for _ in range(batch_size):
        out["flip"].append(rng.random() >= 0.5)
        out["brightness"].append(rng.uniform(0.8, 1.2) if rng.random() >= 0.5 else 1.0)
        out["contrast"].append(rng.uniform(0.8, 1.2) if rng.random() >= 0.5 else 1.0)
        out["noise"].append(rng.random() >= 0.5)
    return out


def hflip_rows_(x, flip):
    """In-place torch.flip(x[b], [-1]) for the samples with flip[b] != 0.  x: [B, ..., W] contiguous, 4- or 8-byte elements."""
    _lib.require_cuda(x, flip)
    if not x.is_contiguous() or x.element_size() not in (4, 8):
        raise ValueError("hflip_rows_: contiguous tensor of 4- or 8-byte elements")
    if flip.dtype not in (torch.uint8, torch.bool) or flip.numel() != x.shape[0]:
        raise ValueError("hflip_rows_: flip must be uint8 / bool [B]")
    B, W = x.shape[0], x.shape[-1]
    rows = x[0].numel() // W if B else 1
    with torch.cuda.device(x.device):
        check(lib().oess_hflip_rows(ptr(x), x.element_size(), B, rows, W, ptr(flip.view(torch.uint8).contiguous()),
                                    stream_ptr(x.device)), "oess_hflip_rows")
    return x


def frame_color_aug_(frame, brightness, contrast, noise=None):
    """In place on frame [B, 3, H, W] float32: adjust_brightness -> adjust_contrast -> + noise (see include/openess_b200.h)."""
    _lib.require_cuda(frame, noise)                            # brightness / contrast: host or device [B] tensors
    if frame.dtype != torch.float32 or not frame.is_contiguous() or frame.ndim != 4 or frame.shape[1] != 3:
        raise ValueError("frame_color_aug_: contiguous float32 [B, 3, H, W]")
    B, _, H, W = frame.shape
    if noise is not None and (noise.shape != frame.shape or noise.dtype != torch.float32 or not noise.is_contiguous()):
        raise ValueError("frame_color_aug_: noise must match frame")
    sums = torch.empty(B, dtype=torch.float64, device=frame.device)
    bf = brightness.to(device=frame.device, dtype=torch.float32).contiguous()
    cf = contrast.to(device=frame.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(frame.device):
        check(lib().oess_frame_color_aug(ptr(frame), B, H * W, ptr(bf), ptr(cf), ptr(noise), ptr(sums), stream_ptr(frame.device)),
              "oess_frame_color_aug")
    return frame


def augment_batch_(event, label, frame, pl, superpixel, params, generator=None):
    """frame2voxel / recon2voxel augmentation of a device batch, in place (`frame` is the frame or the reconstruction).
    params: draw_params(B).  Returns the same tensors."""
    dev = frame.device
    flip = torch.tensor(params["flip"], dtype=torch.uint8).to(dev, non_blocking=True)
    for t in (event, label, frame, pl, superpixel):
        if t is not None:
            hflip_rows_(t, flip)
    noise = None
    if any(params["noise"]):
        gate = torch.tensor(params["noise"], dtype=torch.float32, device=dev).view(-1, 1, 1, 1)
        noise = torch.randn(frame.shape, device=dev, generator=generator) * (0.05 * gate)
    frame_color_aug_(frame, torch.tensor(params["brightness"]), torch.tensor(params["contrast"]), noise)
    return event, label, frame, pl, superpixel
