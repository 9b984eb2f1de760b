#!/usr/bin/env python
"""Secondary measurements (BASELINE.json config 5 + the other kernels of the path), one JSON line each.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.jsonl

* voxelisation throughput sweep, N = 1e4 .. 1e7 events per 640x480 frame, both voxelisers (trilinear =
  VoxelGrid.convert, tbilinear = data_util.generate_voxel_grid), both modes (ordered = bit-exact, atomic),
  uniform and edge-clustered events; DDD17 346x260 frames (BASELINE config 1 voxeliser).
* loss-side kernels at DSEC pretrain shapes (B=8, 440x640, 256-ch features, K=11, S=100).
Every line carries achieved GB/s against the algorithmic bytes of SURVEY.md 8d and the measured HBM peak.
CUDA-event timed, >= 3 warm-ups, inputs larger than L2 or rotated between iterations.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openess_b200 import losses, voxel  # noqa: E402

PEAK = 6539.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    if "bytes" in kw and "ms" in kw:
        kw["gbs"] = kw["bytes"] / (kw["ms"] * 1e-3) / 1e9
        kw["frac_of_hbm_peak"] = kw["gbs"] / PEAK
    print(json.dumps(kw), flush=True)


def synth_xy(rng, n, W, H, clustered):
    if clustered:
        k = int(0.8 * n)
        seg = rng.integers(0, 16, k)
        a = rng.random(k)
        x0, y0, x1, y1 = (rng.uniform(0, s, 16) for s in (W, H, W, H))
        x = np.concatenate([x0[seg] + a * (x1[seg] - x0[seg]) + rng.normal(0, 0.7, k), rng.uniform(0, W, n - k)])
        y = np.concatenate([y0[seg] + a * (y1[seg] - y0[seg]) + rng.normal(0, 0.7, k), rng.uniform(0, H, n - k)])
        p = rng.permutation(n)
        return np.clip(x[p], -0.9, W - 0.1), np.clip(y[p], -0.9, H - 0.1)
    return rng.uniform(-0.9, W - 0.1, n), rng.uniform(-0.9, H - 0.1, n)


def voxel_sweep(dev, quick):
    rng = np.random.default_rng(1205)
    C, H, W = 5, 480, 640
    Ns = [10_000, 100_000, 1_000_000] if quick else [10_000, 30_000, 100_000, 300_000, 1_000_000, 3_000_000, 10_000_000]
    for N in Ns:
        F = int(max(1, min(160, 16_000_000 // N)))
        for clustered in (False, True):
            # one frame's events, tiled F times with different time jitter is enough to defeat caching (F*16N >> L2 for
            # small N is not needed: outputs F*6.1 MB dominate)
            parts = []
            for f in range(min(F, 8)):
                x, y = synth_xy(rng, N, W, H, clustered)
                t = np.sort(rng.integers(0, 50000, N)).astype(np.float64)
                t = (t - t[0]).astype(np.float32)
                t = t / t[-1]
                parts.append((x.astype(np.float32), y.astype(np.float32), rng.integers(0, 2, N).astype(np.float32), t))
            idx = [f % len(parts) for f in range(F)]
            tri = [torch.from_numpy(np.concatenate([parts[i][k] for i in idx])).to(dev) for k in range(4)]
            fo = (torch.arange(F + 1, dtype=torch.int64) * N).to(dev)
            out = torch.empty((F, C, H, W), dtype=torch.float32, device=dev)
            ev4 = torch.stack([tri[0].clamp(min=0).to(torch.int64), tri[1].clamp(min=0).to(torch.int64),
                               (tri[3] * 50000).to(torch.int64) + 1_500_000_000, tri[2].to(torch.int64)], 1).contiguous()
            iters = 5 if N * F >= 8_000_000 else 10
            for mode in ("ordered", "atomic"):
                ms = timeit(lambda: voxel.voxel_trilinear(*tri, C, H, W, frame_offsets=fo, mode=mode, out=out), iters)
                emit(kernel="voxel_trilinear", mode=mode, events_per_frame=N, frames=F, clustered=clustered, ms=ms,
                     frames_per_s=F / (ms * 1e-3), mev_per_s=N * F / (ms * 1e-3) / 1e6, bytes=F * (16 * N + 4 * C * H * W))
                ms = timeit(lambda: voxel.voxel_tbilinear(ev4, C, H, W, frame_offsets=fo, separate_pol=False, mode=mode,
                                                          mutate_p=False, out=out), iters)
                emit(kernel="voxel_tbilinear_i64", mode=mode, events_per_frame=N, frames=F, clustered=clustered, ms=ms,
                     frames_per_s=F / (ms * 1e-3), mev_per_s=N * F / (ms * 1e-3) / 1e6, bytes=F * (32 * N + 4 * C * H * W),
                     bytes_note="32 B/event int64 rows as the reference passes them (14 B native)")
            del tri, ev4, out
    # DDD17 geometry (BASELINE config 1 voxeliser): 346x260, 50k and 32k events, 160 frames per call
    H, W = 260, 346
    for N in (50_000, 32_000):
        F = 160
        x = rng.integers(0, W, N * F)
        y = rng.integers(0, H, N * F)
        t = np.concatenate([np.sort(rng.integers(0, 50000, N)) + 1_500_000_000 for _ in range(F)])
        ev4 = torch.from_numpy(np.stack([x, y, t, rng.integers(0, 2, N * F)], 1).astype(np.int64)).to(dev)
        fo = (torch.arange(F + 1, dtype=torch.int64) * N).to(dev)
        out = torch.empty((F, C, H, W), dtype=torch.float32, device=dev)
        for mode in ("ordered", "atomic"):
            ms = timeit(lambda: voxel.voxel_tbilinear(ev4, C, H, W, frame_offsets=fo, separate_pol=False, mode=mode,
                                                      mutate_p=False, out=out), 10)
            emit(kernel="voxel_tbilinear_i64", geometry="DDD17 346x260", mode=mode, events_per_frame=N, frames=F, ms=ms,
                 frames_per_s=F / (ms * 1e-3), mev_per_s=N * F / (ms * 1e-3) / 1e6, bytes=F * (32 * N + 4 * C * H * W))
        hist = torch.empty((F, 2, H, W), dtype=torch.float32, device=dev)
        ms = timeit(lambda: voxel.voxel_histogram(ev4, H, W, frame_offsets=fo, mutate_p=False, out=hist), 10)
        emit(kernel="voxel_histogram_i64", geometry="DDD17 346x260", events_per_frame=N, frames=F, ms=ms,
             frames_per_s=F / (ms * 1e-3), bytes=F * (32 * N + 8 * H * W))


def loss_sweep(dev, quick):
    g = torch.Generator(device="cpu").manual_seed(1205)
    B, Cf, H, W, K, S = (4 if quick else 8), 256, 440, 640, 11, 100
    feat = torch.randn((B, Cf, H, W), generator=g).to(dev)
    # Voronoi-like superpixels: 10x10 blocks with jittered borders
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    seg = ((yy * 10 // H) * 10 + (xx * 10 // W)).to(torch.int64)
    seg = seg.unsqueeze(0).repeat(B, 1, 1).to(dev)
    M = B * S
    ms = timeit(lambda: losses.segpool_forward(feat, seg, S, M), 10)
    emit(kernel="segpool_fwd", B=B, Cf=Cf, H=H, W=W, S=S, ms=ms, bytes=4 * B * Cf * H * W + 8 * B * H * W + 4 * M * Cf)
    pooled, counts = losses.segpool_forward(feat, seg, S, M)
    dp = torch.randn_like(pooled)
    ms = timeit(lambda: losses.segpool_backward(dp, seg, counts, S, tuple(feat.shape)), 10)
    emit(kernel="segpool_bwd", B=B, Cf=Cf, H=H, W=W, S=S, ms=ms, bytes=4 * B * Cf * H * W + 8 * B * H * W + 4 * M * Cf)
    # the reference's formulation on the same GPU, for scale (torch.sparse one-hot matmul on a permuted copy)
    def ref_pool():
        sp = (torch.arange(0, B * S, S, device=dev)[:, None, None] + seg).flatten()
        oh = torch.sparse_coo_tensor(torch.stack((sp, torch.arange(sp.numel(), device=dev))), torch.ones(sp.numel(), device=dev))
        k = oh @ feat.permute(0, 2, 3, 1).flatten(0, 2)
        return k / (torch.sparse.sum(oh, 1).to_dense()[:, None] + 1e-6)
    try:
        ms = timeit(ref_pool, 3, warm=1)
        emit(kernel="segpool_fwd_torch_sparse_reference_formulation", B=B, ms=ms)
    except Exception as e:  # pragma: no cover
        emit(kernel="segpool_fwd_torch_sparse_reference_formulation", error=str(e)[:100])
    del feat, dp
    for Mn in (800, 3200):
        k = torch.nn.functional.normalize(torch.randn((Mn, 256), generator=g), dim=1).to(dev).requires_grad_(True)
        q = torch.nn.functional.normalize(torch.randn((Mn, 256), generator=g), dim=1).to(dev).requires_grad_(True)
        ms = timeit(lambda: losses._infonce_raw(k, q, 0.07, True), 10)
        emit(kernel="infonce_fwd+bwd", M=Mn, D=256, ms=ms, gflop=3 * 2 * 2 * Mn * Mn * 256 / 1e9 / 2 * 5 / 3,
             tflops_fp32=(5 * 2 * Mn * Mn * 256) / (ms * 1e-3) / 1e12)

        def ref_nce():
            kk, qq = k.detach().requires_grad_(True), q.detach().requires_grad_(True)
            torch.nn.functional.cross_entropy(kk @ qq.T / 0.07, torch.arange(Mn, device=dev)).backward()
        emit(kernel="infonce_fwd+bwd_torch_reference_formulation", M=Mn, ms=timeit(ref_nce, 5))
    logits = torch.randn((B, K, H, W), generator=g).to(dev)
    target = torch.randint(0, K, (B, H, W), generator=g).to(dev)
    target[torch.rand((B, H, W), generator=g).to(dev) < 0.02] = 255
    ms = timeit(lambda: losses.dice_ce_partials(logits, target, 255), 10)
    emit(kernel="dice_ce_partials", B=B, K=K, ms=ms, bytes=B * H * W * (4 * K + 8))
    part = losses.dice_ce_partials(logits, target, 255)
    dl = torch.empty_like(logits)
    gs = torch.ones(1, device=dev)
    from openess_b200._lib import check, lib, ptr, stream_ptr
    ms = timeit(lambda: check(lib().oess_dice_ce_bwd(ptr(logits), ptr(target), B, K, H, W, 255, ptr(part), 1.0, 1.0, ptr(gs),
                                                     ptr(dl), stream_ptr(dev)), "bwd"), 10)
    emit(kernel="dice_ce_bwd", B=B, K=K, ms=ms, bytes=B * H * W * (8 * K + 8))
    pred = torch.randint(0, K, (B, H, W), generator=g).to(dev)
    ms = timeit(lambda: losses.confusion(pred, target, K, 255), 10)
    emit(kernel="confusion", B=B, ms=ms, bytes=16 * B * H * W)
    x = torch.randn((B, 5, H, W), generator=g).to(dev)
    x[torch.rand(x.shape, generator=g).to(dev) < 0.7] = 0
    ms = timeit(lambda: voxel.nonzero_standardize(x, 1, False), 10)
    emit(kernel="nonzero_standardize (EventPreprocessor)", B=B, ms=ms, bytes=12 * x.numel())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    emit(kind="header", peak_hbm_gbs=PEAK, gpu=torch.cuda.get_device_name(0))
    if args.only in ("", "voxel"):
        voxel_sweep(dev, args.quick)
    if args.only in ("", "loss"):
        loss_sweep(dev, args.quick)


if __name__ == "__main__":
    main()
