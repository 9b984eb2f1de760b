#!/usr/bin/env python
"""End-to-end OpenESS pretraining step (BASELINE config 3 per-GPU shard): samples/s and
event-frames/s (= 20 * samples/s) of ONE GPU running the full frame2voxel step at real widths:
raw DSEC records (20 x 100 000 events / sample, pinned host memory) -> H2D -> rectify + voxelise (bit-exact ordered mode)
-> EventPreprocessor + E2VID x 20 (tcgen05) -> SemSegE2VID (cuDNN trunk + fused head) ; frame -> dilated ResNet-50 teacher
(tcgen05, train-mode BN) -> superpixel pooling -> InfoNCE + Dice/CE -> backward -> 2 x AdamW.

    python tools/bench_train_step.py [--batch 4] [--steps 5] [--warmup 2]            # torchrun for N > 1 (gradient all-reduce)
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--events", type=int, default=100_000)
    args = ap.parse_args()
    import bench as vb                                    # synthetic DSEC records + rectify map of the voxel bench
    from seeded_weights import seeded_state_dict
    from openess_b200 import _lib, parallel
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.e2vid.model.model import E2VIDRecurrent
    from openess_b200.models.image_model import DilationFeatureExtractor
    from openess_b200.models.style_networks import SemSegE2VID
    from openess_b200.training.pretrain_step import OpenESSPretrainStep, RawEvents
    from openess_b200.utils.loss_functions import NCELoss, TaskLoss

    rank, local, world = parallel.init()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = True                # torch default: what the reference's own GPU run uses
    torch.backends.cuda.matmul.allow_tf32 = True
    B, Hs, Ws, Hc, K, S, NF = args.batch, 480, 640, 440, 11, 100, 20
    cfg = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
           'base_num_channels': 32, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}
    e2vid = E2VIDRecurrent(cfg, latent_only=True)
    e2vid.load_state_dict(seeded_state_dict(e2vid, 1205), strict=True)
    e2vid = e2vid.eval().to(dev).fold_bn()
    back = SemSegE2VID(input_c=256, output_c=K, skip_connect=True, skip_type='concat', text_embeddings_path='').to(dev)
    teacher = DilationFeatureExtractor()
    teacher.load_state_dict(seeded_state_dict(teacher, 77), strict=True)
    teacher = teacher.to(dev)
    opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
    rec = ImageReconstructor(e2vid, Hc, Ws, 5, dev, opts)
    step = OpenESSPretrainStep(rec, back, teacher, TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255),
                               NCELoss(temperature=0.07), nr_events_data_b=NF, superpixel_size=S, data_parallel=world > 1)

    rng = np.random.default_rng(1205 + rank)
    rmap = torch.from_numpy(vb.synth_rectify_map(rng))
    x, y, t, p = vb.synth_raw_frames(rng, B * NF, n=args.events)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    ev = RawEvents(pin(x), pin(y), pin(t), pin(p), torch.arange(0, (B * NF + 1) * args.events, args.events, dtype=torch.int64),
                   rmap.to(dev), (Hs, Ws), Hc)
    frame = torch.from_numpy(rng.random((B, 3, Hc, Ws)).astype(np.float32)).pin_memory()
    pl = rng.integers(0, K, (B, Hc, Ws))
    pl[rng.random(pl.shape) < 0.02] = 255
    pl = torch.from_numpy(pl.astype(np.int64)).pin_memory()
    # 100 Voronoi-like superpixels per image: nearest of 100 random seeds on a coarse grid
    yy, xx = np.mgrid[0:Hc, 0:Ws]
    sps = []
    for _ in range(B):
        sx, sy = rng.uniform(0, Ws, S), rng.uniform(0, Hc, S)
        d = (xx[None] - sx[:, None, None]) ** 2 + (yy[None] - sy[:, None, None]) ** 2
        sps.append(d.argmin(0))
    sp = torch.from_numpy(np.stack(sps).astype(np.int64)).pin_memory()
    batch = (ev, None, frame, pl, sp)

    def one():
        losses, _, total = step.train_step(batch)
        return float(total.detach())                       # D2H read of the step's loss (the trainer logs it)

    for _ in range(args.warmup):
        last = one()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(args.steps):
        last = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = float(tt)
    if rank == 0:
        print(json.dumps({"metric": "end-to-end pretrain step (frame2voxel), event-frames/s = 20 x samples/s", "n_gpus": world,
                          "batch_per_gpu": B, "ms_per_step": ms, "samples_per_s": world * B / ms * 1e3,
                          "event_frames_per_s": world * B * NF / ms * 1e3, "loss": last,
                          "own_kernel_launches_per_step": (_lib.launch_count() - n0) / args.steps,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                          "h2d_bytes_per_step": int(x.nbytes + y.nbytes + t.nbytes + p.nbytes + frame.numel() * 4 + pl.numel() * 8 + sp.numel() * 8)}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
