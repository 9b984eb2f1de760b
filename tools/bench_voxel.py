#!/usr/bin/env python
"""Single-stream A/B timing of the ordered trilinear voxeliser at the bench workload (F frames of 100 k events, every second
frame edge-clustered), over a list of environment-variable configurations of the kernel plan (read on every call).

    python tools/bench_voxel.py [--frames 160] [--steps 10] CONFIG [CONFIG ...]
    CONFIG = comma-separated KEY=VALUE pairs, e.g.  OESS_TRI_SPLAT=strip   or   OESS_TILE_TH=32,OESS_TILE_THREADS=256

Every configuration's output is compared bit for bit with the first configuration's (use the strip path, validated against the
oracle by the test suite, as the first).  Prints one JSON line per configuration with the per-kernel CUDA-event times."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from openess_b200 import _lib, voxel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=160)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--events", type=int, default=bench.N_EVENTS)
    ap.add_argument("--clustered-every", type=int, default=2)
    ap.add_argument("configs", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    _lib.lib()
    F, C, H, W = args.frames, bench.C, bench.H, bench.W
    n = args.events
    rng = np.random.default_rng(1205)
    rmap = torch.from_numpy(bench.synth_rectify_map(rng)).to(dev)
    fo = (torch.arange(F + 1, dtype=torch.int64) * n).to(dev)
    sets = []
    for _ in range(2):
        raw = bench.synth_raw_frames(rng, F, n=n, clustered_every=args.clustered_every)
        sets.append(list(voxel.dsec_rectify_tnorm(*(torch.from_numpy(a).to(dev) for a in raw), rmap, fo)))
    out = torch.empty((F, C, H, W), dtype=torch.float32, device=dev)
    ref = None
    alg = (16 * n + 4 * C * H * W) * F
    knobs = set()
    for cfg in args.configs:
        for k in knobs:
            os.environ.pop(k, None)
        for kv in cfg.split(","):
            if "=" in kv:
                k, v = kv.split("=", 1)
                os.environ[k] = v
                knobs.add(k)

        def step(i):
            x, y, p, t = sets[i & 1]
            voxel.voxel_trilinear(x, y, p, t, C, H, W, frame_offsets=fo, mode="ordered", out=out)

        step(0)
        torch.cuda.synchronize()
        got = out.clone()
        same = None
        if ref is None:
            ref = got
        else:
            same = bool(torch.equal(got.view(torch.int32), ref.view(torch.int32)))
        del got
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        with _lib.profile() as p:
            for i in range(args.steps):
                step(i)
            torch.cuda.synchronize()
        kern = {k: round(v[1] / args.steps, 4) for k, v in p.kernels.items()}
        print(json.dumps({"config": cfg, "ms_per_step": round(ms, 4), "frames_per_s": round(F / ms * 1e3),
                          "path_frac": round(alg / (ms * 1e-3) / 1e9 / 6545.0, 4), "bit_equal_to_first": same,
                          "kernel_ms": kern}), flush=True)


if __name__ == "__main__":
    main()
