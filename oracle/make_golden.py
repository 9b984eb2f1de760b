"""Generate golden input/output vectors by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests / fixtures for this path, so these vectors are the pin:
they are produced by importing the reference modules by file path (SURVEY.md 7.0 recipe) and calling
them on seeded synthetic inputs that cover the edge cases listed in SURVEY.md 8(a).

* datasets/data_util.py               -> generate_voxel_grid / generate_event_histogram / normalize_voxel_grid
* DSEC/dataset/representations.py     -> VoxelGrid.convert  (torch.set_num_threads(1): serial put_)
* DSEC/dataset/sequence_ov.py         -> events_to_voxel_grid / rectify_events: the module imports h5py
                                         (absent here), so the two method bodies are extracted with `ast`
                                         from the unmodified file and executed as-is.
* utils/loss_functions.py             -> NCELoss, TaskLoss
* evaluation/metrics.py               -> semseg_compute_confusion & co
* training/pretrain_trainer.py:445-465 superpixel pooling is inlined in the trainer; it is re-executed here
                                         with the same torch.sparse calls.
"""
import ast
import hashlib
import importlib.util
import os
import sys
import textwrap

import numpy as np
import torch

REF = os.environ.get("OPENESS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SEED = 1205  # train.py:15-23


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _extract_methods(rel, cls, names):
    """Compile selected methods of a class from the unmodified reference file without importing it."""
    src = open(os.path.join(REF, rel)).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in names:
                    code = textwrap.dedent(ast.get_source_segment(src, fn))
                    ns = {"np": np, "torch": torch}
                    exec(compile(code, rel, "exec"), ns)
                    out[fn.name] = ns[fn.name]
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------------------------------------
def events_i64(rng, n, W, H, span=50000, oob=0.0, pvals=(0, 1), t0=1_500_000_000):
    x = rng.integers(0, W, n)
    y = rng.integers(0, H, n)
    if oob > 0:
        m = rng.random(n) < oob
        x[m] = rng.integers(-3, W + 3, m.sum())
        y[m] = rng.integers(-3, H + 3, m.sum())
    t = np.sort(rng.integers(0, span, n)) + t0
    p = rng.choice(np.array(pvals), n)
    return np.stack([x, y, t, p], 1).astype(np.int64)


def golden_tbilinear(du):
    rng = np.random.default_rng(SEED)
    cases = {}

    def add(name, ev, shape, C, **extra):
        cases[f"{name}__in"] = ev
        cases[f"{name}__meta"] = np.array([shape[0], shape[1], C], np.int64)
        for sp in (False, True):
            e = ev.copy()
            cases[f"{name}__out_sp{int(sp)}"] = du.generate_voxel_grid(e, shape, C, sp)
            cases[f"{name}__pmut"] = e[:, 3].copy()  # polarity column after the call

    add("small", events_i64(rng, 2000, 24, 16, oob=0.1), (16, 24), 5)
    add("n1", events_i64(rng, 1, 24, 16), (16, 24), 5)
    ev = events_i64(rng, 500, 24, 16)
    ev[:, 2] = 77
    add("equal_t", ev, (16, 24), 5)
    ev = events_i64(rng, 20000, 3, 2, span=1000)
    add("hot", ev, (2, 3), 5)
    add("pm1", events_i64(rng, 1500, 24, 16, pvals=(-1, 1)), (16, 24), 5)
    add("pother", events_i64(rng, 1500, 24, 16, pvals=(-3, -1, 0, 1, 2)), (16, 24), 5)
    ev = events_i64(rng, 1500, 24, 16)
    rng.shuffle(ev[:, 2])
    add("unsorted", ev, (16, 24), 5)
    add("bins3", events_i64(rng, 1500, 24, 16), (16, 24), 3)
    add("bins1", events_i64(rng, 300, 24, 16), (16, 24), 1)
    add("bins9", events_i64(rng, 3000, 20, 12), (12, 20), 9)
    # float64 events as produced by np.stack([x_rect, y_rect, t, p]) in sequence_ov.py:268
    n = 2500
    evf = np.stack([rng.uniform(-1.5, 25.5, n).astype(np.float32), rng.uniform(-1.5, 17.5, n).astype(np.float32),
                    np.sort(rng.integers(0, 50000, n)) + 5_000_000, rng.integers(0, 2, n)], 1).astype(np.float64)
    add("f64", evf, (16, 24), 5)
    evf2 = evf.copy()
    evf2[:, 2] = np.sort(rng.uniform(0, 1, n))
    add("f64_fract_t", evf2, (16, 24), 5)
    save("tbilinear", **cases)

    # BASELINE config 1 (DDD17 346x260, 50k events): digest + sparse sample only
    ev = events_i64(rng, 50000, 346, 260)
    out = du.generate_voxel_grid(ev.copy(), (260, 346), 5, False)
    idx = rng.integers(0, out.size, 4096)
    save("tbilinear_ddd17", ev=ev, sha256=np.array(sha(out)), sample_idx=idx, sample_val=out.ravel()[idx],
         meta=np.array([260, 346, 5], np.int64))


def golden_histogram(du):
    rng = np.random.default_rng(SEED + 1)
    cases = {}
    for name, ev, shape in [
        ("small", events_i64(rng, 3000, 24, 16), (16, 24)),
        ("pm1", events_i64(rng, 3000, 24, 16, pvals=(-1, 1)), (16, 24)),
        ("pother", events_i64(rng, 3000, 24, 16, pvals=(-2, 0, 1, 3)), (16, 24)),
        ("hot", events_i64(rng, 50000, 2, 2), (2, 2)),
    ]:
        e = ev.copy()
        cases[f"{name}__in"] = ev
        cases[f"{name}__out"] = du.generate_event_histogram(e, shape)
        cases[f"{name}__pmut"] = e[:, 3].copy()
        cases[f"{name}__meta"] = np.array(shape, np.int64)
    save("histogram", **cases)


def dsec_like(rng, n, W, H, lo=-1.5, hi_pad=0.5, span=50000):
    x = rng.uniform(lo, W + hi_pad, n).astype(np.float32)
    y = rng.uniform(lo, H + hi_pad, n).astype(np.float32)
    pol = rng.integers(0, 2, n).astype(np.float32)
    t = np.sort(rng.integers(0, span, n)).astype(np.float64)
    t = (t - t[0]).astype(np.float32)
    t = t / t[-1]
    return x, y, pol, t


def golden_trilinear(rp):
    torch.set_num_threads(1)  # serial put_
    rng = np.random.default_rng(SEED + 2)
    cases = {}

    def add(name, x, y, pol, t, C, H, W, normalize=False):
        vg = rp.VoxelGrid(C, H, W, normalize)
        with np.errstate(all="ignore"):
            out = vg.convert(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(pol),
                             torch.from_numpy(t)).numpy()
        cases[f"{name}__x"], cases[f"{name}__y"] = x, y
        cases[f"{name}__pol"], cases[f"{name}__t"] = pol, t
        cases[f"{name}__out"] = out
        cases[f"{name}__meta"] = np.array([C, H, W, int(normalize)], np.int64)

    add("small", *dsec_like(rng, 3000, 24, 16), 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 2000, 24, 16)
    x = rng.uniform(-0.999, 1.0, 2000).astype(np.float32)  # x,y in (-1, 1): trunc-toward-zero quirk
    y = rng.uniform(-0.999, 1.0, 2000).astype(np.float32)
    add("negquirk", x, y, pol, t, 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 1, 24, 16)
    add("n1", x, y, pol, np.zeros(1, np.float32) / np.zeros(1, np.float32), 5, 16, 24)  # 0/0 = NaN like :156
    x, y, pol, t = dsec_like(rng, 400, 24, 16)
    add("equal_t", x, y, pol, np.full(400, 0.5, np.float32), 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 30000, 3, 2, lo=0.0, hi_pad=-1.0)
    add("hot", x, y, pol, t, 5, 2, 3)  # >= 1e4 adds per voxel: order sensitivity
    x, y, pol, t = dsec_like(rng, 2500, 24, 16)
    add("raw_t", x, y, pol, (t * 977.0 + 1234.5).astype(np.float32), 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 2500, 24, 16)
    tt = t.copy()
    rng.shuffle(tt)
    add("unsorted_t", x, y, pol, tt, 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 2500, 24, 16)
    x[:5] = [np.nan, np.inf, -np.inf, 3e9, -3e9]
    y[5:8] = [np.nan, 1e20, -1e20]
    add("nonfinite_xy", x, y, pol, t, 5, 16, 24)
    x, y, pol, t = dsec_like(rng, 2500, 24, 16)
    add("intcoords", np.floor(x), np.floor(y), pol, t, 5, 16, 24)
    add("pol_pm1", x, y, 2 * pol - 1, t, 5, 16, 24)
    add("bins3", *dsec_like(rng, 2500, 24, 16), 3, 16, 24)
    add("bins1", *dsec_like(rng, 500, 24, 16), 1, 16, 24)
    add("bins10", *dsec_like(rng, 4000, 20, 12), 10, 12, 20)
    add("normalize", *dsec_like(rng, 3000, 24, 16), 5, 16, 24, normalize=True)
    add("mid", *dsec_like(rng, 20000, 160, 120), 5, 120, 160)
    save("trilinear", **cases)

    # BASELINE config 2 shape (DSEC 640x480, 100k events): digest + sparse sample only
    x, y, pol, t = dsec_like(rng, 100000, 640, 480, lo=-0.75, hi_pad=-0.25)
    out = rp.VoxelGrid(5, 480, 640, False).convert(*(torch.from_numpy(a) for a in (x, y, pol, t))).numpy()
    idx = rng.integers(0, out.size, 4096)
    save("trilinear_dsec", x=x, y=y, pol=pol, t=t, sha256=np.array(sha(out)), sample_idx=idx,
         sample_val=out.ravel()[idx], meta=np.array([5, 480, 640, 0], np.int64))


def golden_dsec_prestep():
    """sequence_ov.py:154-165 (events_to_voxel_grid) + :204-210 (rectify_events), executed from the file."""
    torch.set_num_threads(1)
    m = _extract_methods("DSEC/dataset/sequence_ov.py", "Sequence", {"events_to_voxel_grid", "rectify_events"})
    rng = np.random.default_rng(SEED + 3)
    H, W, C, n = 48, 64, 5, 6000
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rmap = np.stack([xx, yy], -1).astype(np.float32) + rng.uniform(-0.75, 0.75, (H, W, 2)).astype(np.float32)
    x = rng.integers(0, W, n).astype(np.uint16)
    y = rng.integers(0, H, n).astype(np.uint16)
    t = (np.sort(rng.integers(0, 50000, n)) + 51_234_567_890).astype(np.int64)
    p = rng.integers(0, 2, n).astype(np.uint8)

    captured = {}

    class _VG:  # stands in for representations.VoxelGrid to capture what convert() receives
        def convert(self, x, y, pol, t):
            captured.update(x=x.numpy().copy(), y=y.numpy().copy(), pol=pol.numpy().copy(), t=t.numpy().copy())
            return None

    class _Self:
        locations = ["left"]
        height, width = H, W
        rectify_ev_maps = {"left": rmap}
        voxel_grid = _VG()

    s = _Self()
    xy = m["rectify_events"](s, x, y, "left")
    events = np.stack([xy[:, 0], xy[:, 1], t, p], axis=-1)  # sequence_ov.py:303 (float64)
    m["events_to_voxel_grid"](s, events[:, 0], events[:, 1], events[:, 3], events[:, 2])
    save("dsec_prestep", x=x, y=y, t=t, p=p, rectify_map=rmap, xo=captured["x"], yo=captured["y"],
         po=captured["pol"], to=captured["t"], meta=np.array([C, H, W], np.int64))


def golden_normalize(du):
    torch.set_num_threads(1)
    rng = np.random.default_rng(SEED + 4)
    g = rng.normal(0, 1.5, (2, 5, 20, 28)).astype(np.float32)
    g[rng.random(g.shape) < 0.7] = 0
    out = du.normalize_voxel_grid(torch.from_numpy(g.copy())).numpy()
    z = np.zeros((1, 5, 4, 4), np.float32)
    outz = du.normalize_voxel_grid(torch.from_numpy(z.copy())).numpy()
    save("normalize", x=g, out=out, zeros=z, zeros_out=outz)


def golden_losses():
    torch.set_num_threads(1)
    lf = _load("ref_loss_functions", "utils/loss_functions.py")
    mt = _load("ref_metrics", "evaluation/metrics.py")
    rng = np.random.default_rng(SEED + 5)
    torch.manual_seed(SEED)

    # --- superpixel pooling, pretrain_trainer.py:445-465 (same torch.sparse calls) + NCELoss
    B, Cf, H, W, S = 2, 16, 12, 20, 10
    fv = torch.from_numpy(rng.normal(0, 1, (B, Cf, H, W)).astype(np.float32)).requires_grad_(True)
    ff = torch.from_numpy(rng.normal(0, 1, (B, Cf, H, W)).astype(np.float32)).requires_grad_(True)
    sp = torch.from_numpy(rng.integers(0, S, (B, H, W)).astype(np.int64))
    sp[0, :2, :3] = 13  # id >= S aliases into the next sample's range
    superpixels = torch.arange(0, B * S, S)[:, None, None] + sp
    sI = superpixels.flatten()
    total = sI.shape[0]
    with torch.no_grad():
        one_hot = torch.sparse_coo_tensor(torch.stack((sI, torch.arange(total)), 0), torch.ones(total))
    k = one_hot @ fv.permute(0, 2, 3, 1).flatten(0, 2)
    k = k / (torch.sparse.sum(one_hot, 1).to_dense()[:, None] + 1e-6)
    q = one_hot @ ff.permute(0, 2, 3, 1).flatten(0, 2)
    q = q / (torch.sparse.sum(one_hot, 1).to_dense()[:, None] + 1e-6)
    loss = lf.NCELoss(0.07)(k, q)
    loss.backward()
    pool = dict(feat_voxel=fv.detach().numpy(), feat_frame=ff.detach().numpy(), superpixels=sp.numpy(),
                S=np.array(S), k=k.detach().numpy(), q=q.detach().numpy(), nce=np.array(loss.item()),
                d_feat_voxel=fv.grad.numpy(), d_feat_frame=ff.grad.numpy())

    # --- TaskLoss = Dice + CE, ignore 255
    B, K, H, W = 2, 11, 14, 18
    logits = torch.from_numpy(rng.normal(0, 2, (B, K, H, W)).astype(np.float32)).requires_grad_(True)
    target = rng.integers(0, K, (B, H, W)).astype(np.int64)
    target[rng.random(target.shape) < 0.05] = 255
    tl = lf.TaskLoss(losses=["dice", "cross_entropy"], num_classes=K, ignore_index=255)
    # make_one_hot allocates on cuda if available (loss_functions.py:54); CPU here
    lt = tl(logits, torch.from_numpy(target))
    lt.backward()
    dice_only = lf.DiceLoss(num_classes=K, ignore_index=255)(logits.detach(), torch.from_numpy(target))
    task = dict(logits=logits.detach().numpy(), target=target, total=np.array(lt.item()),
                dice=np.array(dice_only.item()), dlogits=logits.grad.numpy())

    # --- metrics
    K = 11
    pred = rng.integers(0, K, (3, 30, 40)).astype(np.int64)
    gt = rng.integers(0, K, (3, 30, 40)).astype(np.int64)
    gt[rng.random(gt.shape) < 0.1] = 255
    gt[gt == 7] = 3  # class 7 absent in gt
    conf = mt.semseg_compute_confusion(torch.from_numpy(pred), torch.from_numpy(gt), K, 255)
    ms = mt.MetricsSemseg(K, 255, [str(i) for i in range(K)])
    ms.update_batch(torch.from_numpy(pred[:2]), torch.from_numpy(gt[:2]))
    ms.update_batch(torch.from_numpy(pred[2:]), torch.from_numpy(gt[2:]))
    summ = ms.get_metrics_summary()
    met = dict(pred=pred, gt=gt, conf=conf.numpy(), miou=np.array(float(summ["miou"])),
               acc=np.array(float(summ["acc"])), iou=np.array([float(summ[str(i)]) for i in range(K)]))
    save("losses", **{f"pool__{k_}": v for k_, v in pool.items()}, **{f"task__{k_}": v for k_, v in task.items()},
         **{f"met__{k_}": v for k_, v in met.items()})


def _extract_functions(rel, names):
    """Compile selected module-level functions of an unmodified reference file without importing it (the module imports
    matplotlib / cv2 at the top)."""
    src = open(os.path.join(REF, rel)).read()
    ns = {"np": np, "os": os}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(textwrap.dedent(ast.get_source_segment(src, node)), rel, "exec"), ns)
    return {n: ns[n] for n in names}


def golden_ddd17_ingest(du):
    """SURVEY 8f row 1 (DDD17 half): the reference's own load_events / extract_events_from_memmap
    (datasets/extract_data_tools/example_loader_ddd17.py:32-54) on a synthetic recording directory, followed by the
    chunking of DDD17Events.__getitem__ (datasets/ddd17_events_loader.py:153-177) and generate_voxel_grid."""
    import tempfile
    fns = _extract_functions("datasets/extract_data_tools/example_loader_ddd17.py", ["load_events", "extract_events_from_memmap"])
    rng = np.random.default_rng(1717)
    n, W, H = 24000, 346, 260
    t = (np.sort(rng.integers(0, 400_000, n)) + 1_500_000_000).astype(np.int64).reshape(n, 1)
    xyp = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(0, 2, n)], 1).astype(np.int16)
    xyp[rng.random(n) < 0.01, 0] = W + 3                       # a few out-of-sensor records
    idx = np.array([[int(t[8000, 0]), 8000, 5000], [int(t[16000, 0]), 16000, 12500], [int(t[23999, 0]), 23999, 20000],
                    [int(t[3000, 0]), 3000, -40]], dtype=np.int64)     # (timestamp, event_idx, event_idx_before)
    out = {"t": t, "xyp": xyp, "index": idx}
    with tempfile.TemporaryDirectory() as d:
        t.tofile(os.path.join(d, "events.dat.t"))
        xyp.tofile(os.path.join(d, "events.dat.xyp"))
        tm, xm = fns["load_events"](os.path.join(d, "events.dat.t"), os.path.join(d, "events.dat.xyp"))
        assert tm.shape == (n, 1) and xm.shape == (n, 3)
        for img_idx in range(4):
            for fixed in (False, True):
                ev = fns["extract_events_from_memmap"](tm, xm, img_idx, idx, fixed, 6000)
                tag = f"s{img_idx}_{int(fixed)}"
                out[tag + "__sha_events"] = np.array(sha(ev))
                out[tag + "__n"] = np.array(ev.shape[0])
                # DDD17Events.__getitem__ :153-177 with nr_events_data = 4, voxel_grid, C = 5, separate_pol False
                nd = 4
                t_ns = ev[:, 2]
                delta = int((t_ns[-1] - t_ns[0]) / nd)
                per = ev.shape[0] // nd
                id_end, grids, cuts = 0, [], [0]
                for i in range(nd):
                    id_start = id_end
                    id_end = int(np.searchsorted(t_ns, t_ns[0] + (i + 1) * delta)) if fixed else id_end + per
                    id_end = min(id_end, ev.shape[0])
                    cuts.append(id_end)
                    grids.append(du.generate_input_representation(ev[id_start:id_end], "voxel_grid", (H, W),
                                                                  nr_temporal_bins=5, separate_pol=False))
                out[tag + "__cuts"] = np.array(cuts)
                out[tag + "__sha_grid"] = np.array(sha(np.concatenate(grids, 0)))
                out[tag + "__sum"] = np.array([float(np.abs(g).astype(np.float64).sum()) for g in grids])
    save("ddd17_ingest", **out)


def synth_dsec_recording(seed=2024, n=60000, span_ms=700, t_offset=5_000_000):
    """Synthetic stand-in for a DSEC events.h5 (dict of arrays with the file's dataset names and dtypes)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.integers(3000, span_ms * 1000, n)).astype(np.uint32)          # first events after ms 3, bursts allowed
    t[20000:20400] = t[20000]                                                     # 400 events on one microsecond
    t = np.sort(t)
    ms_to_idx = np.searchsorted(t, np.arange(span_ms + 1, dtype=np.int64) * 1000, side="left").astype(np.uint64)
    return {"events/x": rng.integers(0, 640, n).astype(np.uint16), "events/y": rng.integers(0, 480, n).astype(np.uint16),
            "events/t": t, "events/p": rng.integers(0, 2, n).astype(np.uint8), "ms_to_idx": ms_to_idx,
            "t_offset": np.array(t_offset, dtype=np.int64)}


def golden_dsec_slicer():
    """SURVEY 8f row 1 (DSEC half): the reference's own EventSlicer (DSEC/utils/eventslicer.py, imported unmodified with
    h5py / hdf5plugin stubbed -- they are only used for the type annotation / codec registration) on a dict stand-in for
    events.h5, and the chunk selection of Sequence.__getitem__ (sequence_ov.py:247-254, 282-305) restated on top of it."""
    import types
    for name in ("h5py", "hdf5plugin"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.File = object
            sys.modules[name] = m
    es = _load("ref_eventslicer", "DSEC/utils/eventslicer.py")
    rec = synth_dsec_recording()
    sl = es.EventSlicer(rec)
    off = int(rec["t_offset"])
    out = {"start": np.array(sl.get_start_time_us()), "final": np.array(sl.get_final_time_us())}
    queries = [(off + 100_000, off + 150_000), (off + 3_000, off + 3_001), (off + 0, off + 10_000), (off + 250_123, off + 250_999),
               (off + 699_000, off + 700_000), (off + 699_500, off + 700_500), (off + int(rec["events/t"][20000]), off + int(rec["events/t"][20000]) + 1)]
    for i, (a, b) in enumerate(queries):
        ev = sl.get_events(a, b)
        out[f"w{i}__q"] = np.array([a, b])
        out[f"w{i}__none"] = np.array(ev is None)
        if ev is not None:
            out[f"w{i}__n"] = np.array(ev["t"].size)
            out[f"w{i}__sha"] = np.array(sha(np.concatenate([ev[k].astype(np.int64) for k in ("x", "y", "t", "p")])))
    fq = [(off + 400_000, 20000), (off + 10_000, 5000), (off + 3_000, 100), (off + 700_000, 1000), (off + 700_001, 1000),
          (off + int(rec["events/t"][20000]), 300), (off + 123_456, 40000)]
    for i, (te, n) in enumerate(fq):
        ev = sl.get_events_fixed_num(te, n)
        out[f"f{i}__q"] = np.array([te, n])
        out[f"f{i}__none"] = np.array(ev is None)
        if ev is not None:
            out[f"f{i}__n"] = np.array(ev["t"].size)
            out[f"f{i}__sha"] = np.array(sha(np.concatenate([ev[k].astype(np.int64) for k in ("x", "y", "t", "p")])))
            # sequence_ov.py:282-305: chunks of nr_events_data = 4
            nd = 4
            per = ev["t"].size // nd
            out[f"f{i}__chunk_t0"] = np.array([int(ev["t"][j * per]) if per else -1 for j in range(nd)])
            out[f"f{i}__per"] = np.array(per)
    # duration mode, :247-254: delta_t_us = 40 000, 4 windows with FLOAT boundaries
    ts_end, delta, nd = off + 300_001, 40_001, 4
    per = delta / nd
    ns = []
    for i in range(nd):
        ev = sl.get_events(ts_end - delta + i * per, ts_end - delta + (i + 1) * per)
        ns.append(ev["t"].size)
        out[f"d{i}__sha"] = np.array(sha(np.concatenate([ev[k].astype(np.int64) for k in ("x", "y", "t", "p")])))
    out["d__n"] = np.array(ns)
    out["d__q"] = np.array([ts_end, delta, nd])
    save("dsec_slicer", **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--dsec-slicer" in sys.argv:
        return golden_dsec_slicer()
    if "--ddd17" in sys.argv:
        return golden_ddd17_ingest(_load("ref_data_util", "datasets/data_util.py"))
    du = _load("ref_data_util", "datasets/data_util.py")
    rp = _load("ref_representations", "DSEC/dataset/representations.py")
    golden_tbilinear(du)
    golden_histogram(du)
    golden_trilinear(rp)
    golden_dsec_prestep()
    golden_normalize(du)
    golden_losses()


if __name__ == "__main__":
    sys.exit(main())
