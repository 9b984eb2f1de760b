"""GPU parity tests of the voxelisation path (run on the B200 box: pytest -m gpu).

Every comparison goes through the C ABI (libopeness_b200.so) via the reference-mirroring Python API and is
checked against (a) the committed golden vectors produced by the reference's own code and (b) the CPU oracle
on fresh seeded inputs.  ORDERED mode must be bit-exact; ATOMIC mode is checked to an absolute tolerance of
2e-5 (the reference's own multi-threaded put_ differs run-to-run by 2.4e-6; float atomics
additionally flush denormals)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import bits_equal, golden_cases, load_golden

pytestmark = pytest.mark.gpu
ATOMIC_ATOL = 2e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from openess_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


# ------------------------------------------------------------------ golden vectors (reference outputs)
def test_tbilinear_golden_bit_exact(dev):
    from openess_b200.datasets import data_util
    z = load_golden("tbilinear")
    for name in golden_cases(z):
        H, W, C = (int(v) for v in z[f"{name}__meta"])
        for sp in (0, 1):
            ev = z[f"{name}__in"].copy()
            out = data_util.generate_input_representation(ev, "voxel_grid", (H, W), C, bool(sp))
            assert out.dtype == np.float32
            assert bits_equal(out, z[f"{name}__out_sp{sp}"]), (name, sp)
            assert np.array_equal(ev[:, 3], z[f"{name}__pmut"]), "polarity mutation side effect"


def test_tbilinear_ddd17_config1_digest(dev):
    from openess_b200.datasets import data_util
    z = load_golden("tbilinear_ddd17")
    H, W, C = (int(v) for v in z["meta"])
    out = data_util.generate_voxel_grid(z["ev"].copy(), (H, W), C, False)
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(z["sha256"])


def test_histogram_golden_bit_exact(dev):
    from openess_b200.datasets import data_util
    z = load_golden("histogram")
    for name in golden_cases(z):
        H, W = (int(v) for v in z[f"{name}__meta"])
        ev = z[f"{name}__in"].copy()
        out = data_util.generate_input_representation(ev, "histogram", (H, W))
        assert bits_equal(out, z[f"{name}__out"]), name
        assert np.array_equal(ev[:, 3], z[f"{name}__pmut"])
    assert data_util.generate_input_representation(ev, "ev_segnet", (H, W)) is None   # data_util.py:11-14


def test_trilinear_golden_bit_exact(dev):
    from openess_b200.DSEC.dataset.representations import VoxelGrid
    z = load_golden("trilinear")
    for name in golden_cases(z):
        C, H, W, norm = (int(v) for v in z[f"{name}__meta"])
        args = [torch.from_numpy(z[f"{name}__{k}"]) for k in ("x", "y", "pol", "t")]
        out = VoxelGrid(C, H, W, bool(norm)).convert(*args)
        assert out.device.type == "cpu" and out.dtype == torch.float32 and tuple(out.shape) == (C, H, W)
        ref = z[f"{name}__out"]
        if norm:
            assert np.array_equal(out.numpy() == 0, ref == 0)
            np.testing.assert_allclose(out.numpy(), ref, rtol=2e-5, atol=2e-6)
        else:
            assert bits_equal(out.numpy(), ref), name
        out_dev = VoxelGrid(C, H, W, bool(norm)).convert(*(a.to(dev) for a in args))
        assert out_dev.device == dev          # returned on pol.device, representations.py:21


def test_trilinear_dsec_config2_digest(dev):
    from openess_b200.DSEC.dataset.representations import VoxelGrid
    z = load_golden("trilinear_dsec")
    C, H, W, _ = (int(v) for v in z["meta"])
    out = VoxelGrid(C, H, W, False).convert(*(torch.from_numpy(z[k]) for k in ("x", "y", "pol", "t"))).numpy()
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(z["sha256"])
    assert bits_equal(out.ravel()[z["sample_idx"]], z["sample_val"])


def test_dsec_prestep_golden_bit_exact(dev):
    from openess_b200 import voxel
    z = load_golden("dsec_prestep")
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    xo, yo, po, to = voxel.dsec_rectify_tnorm(*(torch.from_numpy(z[k]).to(dev) for k in ("x", "y", "t", "p")),
                                              torch.from_numpy(z["rectify_map"]).to(dev), status=st)
    for got, want in ((xo, z["xo"]), (yo, z["yo"]), (po, z["po"]), (to, z["to"])):
        assert bits_equal(got.cpu().numpy(), want)
    assert int(st.item()) == 0


def test_normalize_golden(dev):
    from openess_b200.datasets import data_util
    z = load_golden("normalize")
    out = data_util.normalize_voxel_grid(torch.from_numpy(z["x"])).numpy()
    assert np.array_equal(out == 0, z["out"] == 0)
    np.testing.assert_allclose(out, z["out"], rtol=2e-5, atol=2e-6)
    outz = data_util.normalize_voxel_grid(torch.from_numpy(z["zeros"])).numpy()
    assert bits_equal(outz, z["zeros_out"])


# ------------------------------------------------------------------ oracle on fresh seeded inputs
def _dsec_events(rng, n, W, H, clustered=False):
    if clustered:   # 80 % of the events on a few line segments (edge-like), rest uniform
        k = int(0.8 * n)
        seg = rng.integers(0, 12, k)
        a = rng.uniform(0, 1, k)
        x0, y0 = rng.uniform(0, W, 12), rng.uniform(0, H, 12)
        x1, y1 = rng.uniform(0, W, 12), rng.uniform(0, H, 12)
        x = np.concatenate([x0[seg] + a * (x1[seg] - x0[seg]) + rng.normal(0, 0.7, k), rng.uniform(-1, W, n - k)])
        y = np.concatenate([y0[seg] + a * (y1[seg] - y0[seg]) + rng.normal(0, 0.7, k), rng.uniform(-1, H, n - k)])
        perm = rng.permutation(n)
        x, y = x[perm], y[perm]
    else:
        x, y = rng.uniform(-1.2, W + 0.2, n), rng.uniform(-1.2, H + 0.2, n)
    pol = rng.integers(0, 2, n).astype(np.float32)
    t = np.sort(rng.integers(0, 50000, n)).astype(np.float64)
    t = (t - t[0]).astype(np.float32)
    t = t / t[-1]
    return x.astype(np.float32), y.astype(np.float32), pol, t


@pytest.mark.parametrize("n,H,W,C,clustered", [
    (1000, 33, 47, 5, False), (10000, 120, 160, 5, True), (100000, 480, 640, 5, False),
    (100000, 480, 640, 5, True), (5000, 40, 50, 7, False), (300, 9, 1025, 2, False),
    (3000, 1030, 12, 5, False), (3000, 1030, 12, 3, True), (2000, 30, 3000, 5, False),
    # BASELINE config 5 (sweep up to 1e7 events / frame): dense frames -- long same-cell runs, hot voxels with
    # thousands of sequential adds, strips of thousands of records
    (1000000, 480, 640, 5, False), (1000000, 480, 640, 5, True), (3000000, 480, 640, 5, True),
])
def test_trilinear_vs_oracle(dev, oracle, n, H, W, C, clustered):
    from openess_b200 import voxel
    rng = np.random.default_rng(n + H + C)
    x, y, pol, t = _dsec_events(rng, n, W, H, clustered)
    ref = oracle.voxel_trilinear(x, y, pol, t, C, H, W)
    d = [torch.from_numpy(a).to(dev) for a in (x, y, pol, t)]
    out = voxel.voxel_trilinear(*d, C, H, W, mode="ordered")[0].cpu().numpy()
    assert bits_equal(out, ref)
    out2 = voxel.voxel_trilinear(*d, C, H, W, mode="ordered")[0].cpu().numpy()
    assert bits_equal(out, out2), "ordered mode must be deterministic"
    outa = voxel.voxel_trilinear(*d, C, H, W, mode="atomic")[0].cpu().numpy()
    # float atomics add in arrival order: the noise scales with the magnitude of hot voxels (thousands of adds)
    np.testing.assert_allclose(outa, ref, rtol=1e-5, atol=ATOMIC_ATOL)


def test_trilinear_batched_ragged_frames(dev, oracle):
    """F frames per launch with ragged sizes incl. empty and 1-event frames == per-frame oracle."""
    from openess_b200 import voxel
    rng = np.random.default_rng(7)
    H, W, C = 60, 80, 5
    sizes = [0, 1, 2047, 2048, 2049, 0, 5000, 31, 12345, 0]
    parts = [_dsec_events(rng, max(n, 2), W, H) for n in sizes]
    parts = [tuple(a[:n] for a in p) for p, n in zip(parts, sizes)]
    cat = [np.concatenate([p[k] for p in parts]) for k in range(4)]
    fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    d = [torch.from_numpy(a).to(dev) for a in cat]
    for mode in ("ordered", "atomic"):
        out = voxel.voxel_trilinear(*d, C, H, W, frame_offsets=torch.from_numpy(fo), mode=mode).cpu().numpy()
        assert out.shape == (len(sizes), C, H, W)
        for f, n in enumerate(sizes):
            if n == 0:
                assert not out[f].any()
                continue
            with np.errstate(all="ignore"):
                ref = oracle.voxel_trilinear(*(p for p in parts[f]), C, H, W)
            if mode == "ordered":
                assert bits_equal(out[f], ref), f
            else:
                np.testing.assert_allclose(out[f], ref, rtol=0, atol=ATOMIC_ATOL)


def _ddd17_events(rng, n, W, H, oob=0.02):
    x = rng.integers(0, W, n)
    y = rng.integers(0, H, n)
    m = rng.random(n) < oob
    x[m] = rng.integers(-2, W + 2, m.sum())
    y[m] = rng.integers(-2, H + 2, m.sum())
    t = np.sort(rng.integers(0, 50000, n)) + 1_500_000_000
    p = rng.integers(0, 2, n)
    return np.stack([x, y, t, p], 1).astype(np.int64)


@pytest.mark.parametrize("n,H,W,C", [(50000, 260, 346, 5), (32000, 260, 346, 5), (4000, 31, 45, 3),
                                     (100000, 480, 640, 5), (20000, 8, 8, 5)])
def test_tbilinear_vs_oracle(dev, oracle, n, H, W, C):
    from openess_b200 import voxel
    rng = np.random.default_rng(n + W)
    ev = _ddd17_events(rng, n, W, H)
    for sp in (False, True):
        ref = oracle.voxel_tbilinear(ev.copy(), (H, W), C, sp)
        d = torch.from_numpy(ev.copy()).to(dev)
        out = voxel.voxel_tbilinear(d, C, H, W, separate_pol=sp, mode="ordered", mutate_p=True)[0].cpu().numpy()
        assert bits_equal(out, ref)
        p = d[:, 3].cpu().numpy()
        assert np.array_equal(p, np.where(ev[:, 3] == 0, -1, ev[:, 3])), "device-side polarity mutation"
        outa = voxel.voxel_tbilinear(torch.from_numpy(ev.copy()).to(dev), C, H, W, separate_pol=sp,
                                     mode="atomic")[0].cpu().numpy()
        np.testing.assert_allclose(outa, ref, rtol=0, atol=2e-4 if H * W <= 64 else ATOMIC_ATOL)


def test_tbilinear_batched_f64(dev, oracle):
    from openess_b200 import voxel
    rng = np.random.default_rng(11)
    H, W, C = 24, 40, 5
    sizes = [3000, 0, 1, 2500]
    frames = []
    for n in sizes:
        e = _ddd17_events(rng, max(n, 1), W, H).astype(np.float64)[:n]
        e[:, 0] += rng.uniform(0, 0.99, n)
        frames.append(e)
    ev = np.concatenate(frames)
    fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    out = voxel.voxel_tbilinear(torch.from_numpy(ev.copy()).to(dev), C, H, W, frame_offsets=torch.from_numpy(fo),
                                separate_pol=False, mode="ordered").cpu().numpy()
    for f, n in enumerate(sizes):
        if n == 0:
            assert not out[f].any()
        else:
            assert bits_equal(out[f], oracle.voxel_tbilinear(frames[f].copy(), (H, W), C, False)), f


def test_histogram_vs_oracle_and_status(dev, oracle):
    from openess_b200 import voxel
    from openess_b200.datasets import data_util
    rng = np.random.default_rng(3)
    ev = _ddd17_events(rng, 30000, 64, 48, oob=0.0)
    ref = oracle.histogram(ev.copy(), (48, 64))
    out = voxel.voxel_histogram(torch.from_numpy(ev.copy()).to(dev), 48, 64)[0].cpu().numpy()
    assert bits_equal(out, ref)
    bad = ev.copy()
    bad[5, 1] = 48          # flat index beyond H*W -> the reference raises IndexError
    with pytest.raises(IndexError):
        data_util.generate_event_histogram(bad, (48, 64))


def test_reference_error_behaviour(dev):
    from openess_b200.datasets import data_util
    from openess_b200.DSEC.dataset.representations import VoxelGrid
    with pytest.raises(IndexError):
        data_util.generate_voxel_grid(np.zeros((0, 4), np.int64), (4, 4), 5)
    with pytest.raises(AssertionError):
        data_util.generate_voxel_grid(np.zeros((3, 3), np.int64), (4, 4), 5)
    with pytest.raises(IndexError):
        VoxelGrid(5, 4, 4, False).convert(*(torch.zeros(0) for _ in range(4)))
    with pytest.raises(AssertionError):
        VoxelGrid(5, 4, 4, False).convert(torch.zeros(3), torch.zeros(3), torch.zeros(3), torch.zeros(2))


# ------------------------------------------------------------------ full-size, size-independent properties
def test_fullsize_properties(dev, oracle):
    """BASELINE config 2 size, F = 40 frames per launch: checksum of checksums vs float64 weight sums,
    batch == single-frame, ordered == atomic within tolerance, ordered deterministic."""
    from openess_b200 import voxel
    rng = np.random.default_rng(2024)
    H, W, C, F, n = 480, 640, 5, 40, 100000
    parts = [_dsec_events(rng, n, W, H, clustered=bool(f % 2)) for f in range(F)]
    cat = [torch.from_numpy(np.concatenate([p[k] for p in parts])).to(dev) for k in range(4)]
    fo = torch.arange(F + 1, dtype=torch.int64) * n
    out = voxel.voxel_trilinear(*cat, C, H, W, frame_offsets=fo, mode="ordered")
    out_a = voxel.voxel_trilinear(*cat, C, H, W, frame_offsets=fo, mode="atomic")
    assert torch.equal(out, voxel.voxel_trilinear(*cat, C, H, W, frame_offsets=fo, mode="ordered"))
    assert float((out - out_a).abs().max()) <= ATOMIC_ATOL
    for f in (0, 17, F - 1):   # batch element == single-frame call == oracle
        single = voxel.voxel_trilinear(*(a[f * n:(f + 1) * n].contiguous() for a in cat), C, H, W, mode="ordered")[0]
        assert torch.equal(single, out[f])
        assert bits_equal(out[f].cpu().numpy(), oracle.voxel_trilinear(*parts[f], C, H, W))
    # per-frame checksum: sum of the grid == sum of all in-bounds corner weights (float64 reference)
    sums = out.double().sum(dim=(1, 2, 3)).cpu().numpy()
    for f in range(F):
        x, y, pol, t = (a.astype(np.float64) for a in parts[f])
        tn = (C - 1) * t
        tot = 0.0
        x0, y0, t0 = np.trunc(x), np.trunc(y), np.trunc(tn)
        for dx in (0, 1):
            for dy in (0, 1):
                for dt in (0, 1):
                    xl, yl, tl = x0 + dx, y0 + dy, t0 + dt
                    m = (xl >= 0) & (xl < W) & (yl >= 0) & (yl < H) & (tl >= 0) & (tl < C)
                    w = (2 * pol - 1) * (1 - np.abs(xl - x)) * (1 - np.abs(yl - y)) * (1 - np.abs(tl - tn))
                    tot += w[m].sum()
        assert abs(sums[f] - tot) < 2e-2, (f, sums[f], tot)


def test_raw_dsec_records_to_voxel_grid(dev, oracle):
    """GPU-side sample assembly: raw (u16 x, u16 y, u32|i64 t, u8 p) records of F frames -> rectify -> t-normalise
    -> trilinear voxel grid == the oracle's per-frame pipeline, bit for bit."""
    from openess_b200 import voxel
    rng = np.random.default_rng(77)
    H, W, C = 96, 128, 5
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rmap = (np.stack([xx, yy], -1) + rng.uniform(-0.75, 0.75, (H, W, 2))).astype(np.float32)
    sizes = [4000, 1, 2500, 0, 9000]
    xs = [rng.integers(0, W, n).astype(np.uint16) for n in sizes]
    ys = [rng.integers(0, H, n).astype(np.uint16) for n in sizes]
    ts = [(np.sort(rng.integers(0, 50000, n)) + 3_000_000_000 + 60000 * i).astype(np.uint32) for i, n in enumerate(sizes)]
    ps = [rng.integers(0, 2, n).astype(np.uint8) for n in sizes]
    fo = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    cat = lambda parts: torch.from_numpy(np.concatenate(parts)).to(dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    for tdt in (np.uint32, np.int64):
        tcat = torch.from_numpy(np.concatenate(ts).astype(tdt) + (0 if tdt == np.uint32 else 10**12)).to(dev)
        out = voxel.dsec_events_to_voxel_grid(cat(xs), cat(ys), tcat, cat(ps), torch.from_numpy(rmap).to(dev), C,
                                              frame_offsets=fo, status=st).cpu().numpy()
        assert int(st.item()) == 0
        for f, n in enumerate(sizes):
            if n == 0:
                assert not out[f].any()
                continue
            with np.errstate(all="ignore"):
                xo, yo, po, to = oracle.dsec_rectify_tnorm(xs[f], ys[f], ts[f].astype(np.int64), ps[f], rmap)
                ref = oracle.voxel_trilinear(xo, yo, po, to, C, H, W)
            assert bits_equal(out[f], ref), (f, tdt)
    bad = cat(xs).clone()
    bad[3] = W                                                     # sequence_ov.py:208 assert x.max() < width
    voxel.dsec_rectify_tnorm(bad, cat(ys), torch.from_numpy(np.concatenate(ts)).to(dev), cat(ps),
                             torch.from_numpy(rmap).to(dev), fo, status=st)
    assert int(st.item()) == 1
