"""Synthetic DSEC-shaped inputs (seed 1205) shared by bench.py, the tools and the tests."""
import numpy as np

C, H, W = 5, 480, 640
N_EVENTS = 100_000
WINDOW_US = 50_000


def synth_rectify_map(rng):
    """identity + U(-0.75, 0.75) px jitter: ~0.2 % of corners leave the sensor, some x' < 0."""
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return (np.stack([xx, yy], -1) + rng.uniform(-0.75, 0.75, (H, W, 2))).astype(np.float32)


def synth_raw_frames(rng, F, n=N_EVENTS, clustered_every=2):
    """Raw DSEC records for F frames of n events over 50 ms; every `clustered_every`-th frame has 80 % of its events
    on 16 line segments (edge-like, stresses same-voxel accumulation)."""
    xs, ys, ts, ps = [], [], [], []
    for f in range(F):
        if clustered_every and f % clustered_every == clustered_every - 1:
            k = int(0.8 * n)
            seg = rng.integers(0, 16, k)
            a = rng.random(k)
            x0, y0, x1, y1 = (rng.uniform(0, s, 16) for s in (W, H, W, H))
            x = np.concatenate([x0[seg] + a * (x1[seg] - x0[seg]) + rng.normal(0, 0.7, k), rng.uniform(0, W, n - k)])
            y = np.concatenate([y0[seg] + a * (y1[seg] - y0[seg]) + rng.normal(0, 0.7, k), rng.uniform(0, H, n - k)])
            perm = rng.permutation(n)
            x, y = x[perm], y[perm]
        else:
            x, y = rng.uniform(0, W, n), rng.uniform(0, H, n)
        xs.append(np.clip(x, 0, W - 1).astype(np.uint16))
        ys.append(np.clip(y, 0, H - 1).astype(np.uint16))
        ts.append((np.sort(rng.integers(0, WINDOW_US, n)) + 1_000_000 + f * WINDOW_US).astype(np.uint32))
        ps.append(rng.integers(0, 2, n).astype(np.uint8))
    return [np.concatenate(a) for a in (xs, ys, ts, ps)]


def synth_superpixels(rng, B, Hc, Wc, S=100):
    """S Voronoi cells per image (ids 0 .. S-1), int64 [B, Hc, Wc]."""
    yy, xx = np.mgrid[0:Hc, 0:Wc]
    sps = []
    for _ in range(B):
        sx, sy = rng.uniform(0, Wc, S), rng.uniform(0, Hc, S)
        d = (xx[None] - sx[:, None, None]) ** 2 + (yy[None] - sy[:, None, None]) ** 2
        sps.append(d.argmin(0))
    return np.stack(sps).astype(np.int64)
