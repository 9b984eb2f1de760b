#!/usr/bin/env python
"""BASELINE.json configs 1, 4 and 5 as measured configurations (VERDICT r01 "missing" #5), one JSON line each.

    python tools/bench_configs.py --config1                 # CPU: DDD17 50 k-event voxel grid + linear-probe forward (host cores)
    python tools/bench_configs.py --config1 --gpu           # the same forward on the B200 (own kernels) next to it
    [torchrun --nproc-per-node N] python tools/bench_configs.py --config4      # DeepLabv3 head fine-tune step, N GPUs, NCCL
    [torchrun --nproc-per-node N] python tools/bench_configs.py --config5      # voxelisation sweep, frames sharded over N GPUs

Config 1: DDD17 346 x 260, 50 000 events -> generate_voxel_grid (5 bins) ; linear-probe forward =
E2VIDRecurrent (20 recurrent steps) + SemSegE2VID (K = 6, if_linear_probing) at 200 x 352, random weights, fp32, on the CPU.
The reference's modules cannot travel to the GPU box; the CPU arm times this repository's module mirrors in their torch
formulation (state_dict-compatible with the reference classes, pinned to them by tests/golden) and the C oracle voxeliser."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def _linear_probe_modules(dev):
    from seeded_weights import seeded_state_dict
    from openess_b200.e2vid.model.model import E2VIDRecurrent
    from openess_b200.models.style_networks import SemSegE2VID
    from openess_b200.training.bench_step import E2VID_CFG
    e2vid = E2VIDRecurrent(E2VID_CFG, latent_only=True)
    e2vid.load_state_dict(seeded_state_dict(e2vid, 1205), strict=True)
    e2vid = e2vid.eval().to(dev)
    torch.manual_seed(1205)
    back = SemSegE2VID(input_c=256, output_c=6, skip_connect=True, skip_type='concat', text_embeddings_path='',
                       if_linear_probing=True).eval().to(dev)
    return e2vid, back


def config1(gpu):
    from oracle import oracle as orc
    orc.build()
    rng = np.random.default_rng(1205)
    H, W, n = 260, 346, 50_000
    ev = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.integers(0, 50000, n)) + 1_500_000_000,
                   rng.integers(0, 2, n)], 1).astype(np.int64)
    orc.voxel_tbilinear(ev.copy(), (H, W), 5, False)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        grid = orc.voxel_tbilinear(ev.copy(), (H, W), 5, False)
    vox_ms = (time.perf_counter() - t0) / reps * 1e3
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cpu")
    e2vid, back = _linear_probe_modules(dev)
    x = torch.from_numpy(rng.normal(0, 1, (1, 100, 200, 352)).astype(np.float32))
    x[torch.rand(x.shape) < 0.6] = 0

    def forward(e2vid, back, x):
        from types import SimpleNamespace
        from openess_b200.e2vid.image_reconstructor import ImageReconstructor
        if x.is_cuda:
            opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
            rec = ImageReconstructor(e2vid, 200, 352, 5, x.device, opts)
            for i in range(20):
                _, _, latent = rec.update_reconstruction(x[:, 5 * i:5 * i + 5])
        else:
            states = None
            for i in range(20):                               # linear_probe_trainer.py:452-462 / inference_utils.py:77-85
                e = x[:, 5 * i:5 * i + 5]
                nz = e != 0
                cnt = nz.sum()
                mean = e.sum() / cnt
                std = torch.sqrt((e ** 2).sum() / cnt - mean ** 2)
                _, states, latent = e2vid(nz.float() * (e - mean) / std, states)
        pred, _ = back(latent)
        return pred[1]

    with torch.no_grad():
        forward(e2vid, back, x)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            out = forward(e2vid, back, x)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
    line = {"config": "1: DDD17 346x260 single frame, 50k-event voxel grid, linear-probe forward on CPU", "host_cores": os.cpu_count(),
            "voxel_grid_ms_1_thread_oracle_port": vox_ms, "voxel_frames_per_s_1_thread": 1e3 / vox_ms,
            "linear_probe_forward_ms_cpu": cpu_ms, "forward": "E2VIDRecurrent x 20 steps + SemSegE2VID (K=6, linear probe) at 1x100x200x352, fp32, "
            "module mirrors in their torch formulation", "logits_shape": list(out.shape), "voxel_checksum": float(grid.sum())}
    if gpu:
        from openess_b200 import voxel
        d = torch.device("cuda", 0)
        e2g, bg = _linear_probe_modules(d)
        e2g.fold_bn()
        xg = x.to(d)
        evg = torch.from_numpy(ev).to(d)
        with torch.no_grad():
            for _ in range(3):
                forward(e2g, bg, xg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                og = forward(e2g, bg, xg)
            e1.record()
            torch.cuda.synchronize()
            line["linear_probe_forward_ms_b200"] = e0.elapsed_time(e1) / 10
            line["argmax_agreement_cpu_vs_b200"] = float((og.argmax(1).cpu() == out.argmax(1)).float().mean())
            for _ in range(3):
                g = voxel.voxel_tbilinear(evg, 5, H, W, separate_pol=False, mutate_p=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                g = voxel.voxel_tbilinear(evg, 5, H, W, separate_pol=False, mutate_p=False)
            e1.record()
            torch.cuda.synchronize()
            line["voxel_grid_ms_b200_single_frame_call"] = e0.elapsed_time(e1) / 20
            line["voxel_bit_equal_cpu_vs_b200"] = bool(np.array_equal(g[0].cpu().numpy(), grid))
    emit(**line)


def _dist():
    from openess_b200 import parallel
    rank, local, world = parallel.init()
    torch.cuda.set_device(local)
    return rank, local, world


def _timed(fn, iters, warm, world):
    import torch.distributed as dist
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def config4():
    """DeepLabv3-R50 head fine-tune step (frozen backbone, K = 11, 440 x 640, batch 4 per GPU): Dice + CE on labels, AdamW,
    gradients of the 17.9 M head parameters (71.6 MB) all-reduced over NCCL, overlapped with backward."""
    from seeded_weights import seeded_state_dict
    from openess_b200 import parallel
    from openess_b200.models import deeplabv3 as dl
    from openess_b200.utils.loss_functions import TaskLoss
    rank, local, world = _dist()
    dev = torch.device("cuda", local)
    m = dl.deeplabv3_resnet50(num_classes=11, text_embeddings_path=None, output_stride=32, pretrained_backbone='',
                              if_finetuning=True, frozen_backbone=True)
    m.load_state_dict(seeded_state_dict(m, 4), strict=True)
    m = m.to(dev).train()
    B = 4
    g = torch.Generator().manual_seed(rank)
    x = torch.rand(B, 3, 440, 640, generator=g).to(dev)
    y = torch.randint(0, 11, (B, 440, 640), generator=g).to(dev)
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-4, fused=True)
    loss_fn = TaskLoss(losses=['dice', 'cross_entropy'], num_classes=11, ignore_index=255)
    red = parallel.GradientReducer(params) if world > 1 else None

    def step():
        opt.zero_grad(set_to_none=True)
        logits, _ = m(x)
        loss = loss_fn(logits, y)
        if red:
            red.prepare()
        loss.backward()
        if red:
            red.finish()
        opt.step()

    ms = _timed(step, 8, 3, world)
    if rank == 0:
        emit(config="4: DeepLabv3 head fine-tune on DSEC-Semantic 11-class (frozen backbone), batch 4 per GPU", n_gpus=world,
             ms_per_step=ms, samples_per_s=world * B / ms * 1e3, allreduce_bytes=(red.stats["bytes"] if red else 0),
             allreduce_buckets=(red.stats["buckets"] if red else 0))
    if world > 1:
        torch.distributed.destroy_process_group()


def config5(quick):
    """Voxelisation throughput sweep, N events per 640 x 480 frame, ordered / atomic, uniform / clustered: the frames of a
    launch are sharded over the ranks (weak scaling: F frames per GPU), aggregate frames/s = world * F / max-over-ranks time."""
    from openess_b200 import voxel
    from tools_sweep import synth_xy                       # noqa: F401  (set up below)
    rank, local, world = _dist()
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(1205 + rank)
    C, H, W = 5, 480, 640
    Ns = [10_000, 100_000, 1_000_000] if quick else [10_000, 30_000, 100_000, 300_000, 1_000_000, 3_000_000, 10_000_000]
    for N in Ns:
        F = int(max(1, min(160, 16_000_000 // N)))
        for clustered in (False, True):
            parts = []
            for f in range(min(F, 8)):
                x, y = synth_xy(rng, N, W, H, clustered)
                t = np.sort(rng.integers(0, 50000, N)).astype(np.float64)
                t = (t - t[0]).astype(np.float32)
                parts.append((x.astype(np.float32), y.astype(np.float32), rng.integers(0, 2, N).astype(np.float32), t / t[-1]))
            idx = [f % len(parts) for f in range(F)]
            tri = [torch.from_numpy(np.concatenate([parts[i][k] for i in idx])).to(dev) for k in range(4)]
            fo = (torch.arange(F + 1, dtype=torch.int64) * N).to(dev)
            out = torch.empty((F, C, H, W), dtype=torch.float32, device=dev)
            for mode in ("ordered", "atomic"):
                ms = _timed(lambda: voxel.voxel_trilinear(*tri, C, H, W, frame_offsets=fo, mode=mode, out=out),
                            5 if N * F >= 8_000_000 else 10, 3, world)
                if rank == 0:
                    emit(config="5: voxelisation sweep", kernel="voxel_trilinear", mode=mode, events_per_frame=N, frames_per_gpu=F,
                         clustered=clustered, n_gpus=world, ms=ms, frames_per_s=world * F / (ms * 1e-3),
                         frac_of_hbm_peak_per_gpu=F * (16 * N + 4 * C * H * W) / (ms * 1e-3) / 1e9 / 6545.0)
            del tri, out
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config1", action="store_true")
    ap.add_argument("--config4", action="store_true")
    ap.add_argument("--config5", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if args.config1:
        config1(args.gpu)
    if args.config4:
        config4()
    if args.config5:
        import importlib.util
        spec = importlib.util.spec_from_file_location("tools_sweep", os.path.join(ROOT, "tools", "sweep.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["tools_sweep"] = mod
        spec.loader.exec_module(mod)
        config5(args.quick)


if __name__ == "__main__":
    main()
