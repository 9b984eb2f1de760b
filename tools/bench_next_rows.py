#!/usr/bin/env python
"""Measurements for the SURVEY 8f rows: DDD17 record ingest (native 14 B / event records vs the int64 rows the reference
assembles), the GPU augmentation pass, and the E2VID step with online reconstruction.  One JSON line each."""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from openess_b200 import voxel  # noqa: E402
from tools.bench_tc import timeit  # noqa: E402


def ddd17():
    """BASELINE config 1 geometry: 346 x 260, 5 chunks x 32 000 events per sample (ddd17_events_loader.py:40-48), B = 8 samples."""
    from openess_b200.datasets.extract_data_tools import example_loader_ddd17 as ld
    rng = np.random.default_rng(17)
    B, nd, per, H, W = 8, 5, 32000, 260, 346
    n = B * nd * per + 1000
    t = (np.sort(rng.integers(0, 60_000_000, n)) + 1_500_000_000).astype(np.int64).reshape(n, 1)
    xyp = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(0, 2, n)], 1).astype(np.int16)
    with tempfile.TemporaryDirectory() as d:
        t.tofile(os.path.join(d, "events.dat.t"))
        xyp.tofile(os.path.join(d, "events.dat.xyp"))
        t_ev, xyp_ev = ld.load_events(os.path.join(d, "events.dat.t"), os.path.join(d, "events.dat.xyp"))
        index = np.array([[0, (i + 1) * nd * per + 500, 0] for i in range(B)], dtype=np.int64)
        stager = ld.DDD17Stager(B * nd * per)

        def native():          # memmap -> pinned -> H2D (14 B / event) -> one launch
            return ld.load_event_tensors(stager, t_ev, xyp_ev, range(B), index, (H, W), nr_events_data=nd, nr_events=nd * per,
                                         separate_pol=False)

        rows_pinned = torch.empty((B * nd * per, 4), dtype=torch.int64).pin_memory()
        fo = torch.arange(B * nd + 1, dtype=torch.int64) * per

        def reference_rows():  # the reference's host assembly (int64 [n, 4], 32 B / event) -> pinned -> H2D -> one launch
            o = 0
            for i in range(B):
                ev = ld.extract_events_from_memmap(t_ev, xyp_ev, i, index, False, nd * per)
                rows_pinned[o:o + ev.shape[0]] = torch.from_numpy(np.ascontiguousarray(ev))
                o += ev.shape[0]
            return voxel.voxel_tbilinear(rows_pinned.cuda(non_blocking=True), 5, H, W, frame_offsets=fo, separate_pol=False,
                                         mutate_p=False)

        a, b = native(), reference_rows()
        assert torch.equal(a.view(-1), b.view(-1))
        t_nat = timeit(native, iters=10, warm=3)
        t_rows = timeit(reference_rows, iters=5, warm=2)
        td, xd = stager.t_dev[:B * nd * per], stager.xyp_dev[:B * nd * per]
        t_dev = timeit(lambda: voxel.voxel_tbilinear_ddd17(td, xd, 5, H, W, frame_offsets=fo, separate_pol=False), iters=20, warm=3)
    F = B * nd
    print(json.dumps({"op": "ddd17_ingest_voxelise", "frames": F, "events_per_frame": per, "geometry": [H, W],
                      "native_records_ms": round(t_nat, 3), "native_frames_per_s": round(F / t_nat * 1e3),
                      "int64_rows_ms": round(t_rows, 3), "int64_rows_frames_per_s": round(F / t_rows * 1e3),
                      "h2d_bytes_per_event": {"native": 14, "int64_rows": 32},
                      "device_only_ms": round(t_dev, 3), "device_only_frames_per_s": round(F / t_dev * 1e3),
                      "device_algorithmic_gbs": round((14 * per + 4 * 5 * H * W) * F / t_dev / 1e6, 1)}))


def augmentation():
    from openess_b200.DSEC.dataset.augment import augment_batch_
    B, H, W = 8, 440, 640
    event = torch.randn(B, 100, H, W, device="cuda")
    frame = torch.rand(B, 3, H, W, device="cuda")
    label = torch.randint(0, 11, (B, H, W), device="cuda")
    pl = torch.randint(0, 11, (B, H, W), device="cuda")
    sp = torch.randint(0, 100, (B, H, W), device="cuda")
    params = {"flip": [True] * B, "brightness": [1.1] * B, "contrast": [0.9] * B, "noise": [True] * B}
    t_all = timeit(lambda: augment_batch_(event, label, frame, pl, sp, params), iters=10, warm=3)
    byt = 2 * (event.numel() * 4 + frame.numel() * 4 + 3 * label.numel() * 8) + 4 * frame.numel() * 4
    print(json.dumps({"op": "augment_batch_frame2voxel", "B": B, "all_samples_flipped": True, "ms": round(t_all, 3),
                      "algorithmic_gbs": round(byt / t_all / 1e6, 1),
                      "note": "flip of the [B,100,440,640] event tensor dominates (986 MB read + written); noise drawn with torch.randn"}))


def reconstruction():
    from seeded_weights import seeded_state_dict
    from openess_b200.e2vid.model import model as mm
    cfg = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
           'base_num_channels': 32, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}
    m = mm.E2VIDRecurrent(cfg, latent_only=False)
    m.load_state_dict(seeded_state_dict(m, 1205), strict=True)
    m = m.eval().cuda().fold_bn()
    x = torch.randn(8, 5, 440, 640, device="cuda")
    res = {}
    with torch.no_grad():
        _, st, _ = m(x, None)
        for use_tc in (True, False):
            mm.USE_TENSOR_CORES = use_tc
            res[use_tc] = timeit(lambda: m(x, st), iters=10, warm=3)
        mm.USE_TENSOR_CORES = True
        m.latent_only = True
        t_lat = timeit(lambda: m(x, st), iters=10, warm=3)
    print(json.dumps({"op": "e2vid_step_with_online_reconstruction", "B": 8, "H": 440, "W": 640, "ms_own_kernels": round(res[True], 3),
                      "ms_torch_cudnn_tf32": round(res[False], 3), "ms_latent_only_own_kernels": round(t_lat, 3)}))


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    ddd17()
    augmentation()
    reconstruction()
