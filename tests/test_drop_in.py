"""Zero-edit drop-in (openess_b200/training/drop_in.py, installed by patch_reference).

CPU, build container only (needs /root/reference): the REFERENCE's own `Sequence` class (DSEC/dataset/sequence_ov.py, h5py /
hdf5plugin stubbed, an in-memory "h5" behind the event slicer, PNG targets in a temp dir) is served by the reference's stock
`torch.utils.data.DataLoader(num_workers=1)` call: the forked worker returns the raw record slab instead of voxelising, the
registered collate builds a `RawEvents` batch in the main process, and the CPU oracle's voxelisation of that batch equals the
dense tensor the UNPATCHED `__getitem__` computes with the reference's own VoxelGrid -- bit for bit, flips included.

GPU (no reference tree needed): the rebound trainer methods on a trainer object shaped like `OpenESSPretrainModel`."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER_CODE = r'''
import os, sys, types, random
import numpy as np, torch
ROOT, TMP = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
for name in ("h5py", "hdf5plugin", "matplotlib", "matplotlib.pyplot", "albumentations"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["h5py"].File = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib.pyplot"].cm = types.SimpleNamespace(Blues=None)
from PIL import Image
from pathlib import Path

def build_sequence(Sequence, EventSlicer, VoxelGrid, rng):
    H, W, nd, per = 480, 640, 4, 3000
    n = 40000
    t = np.sort(rng.integers(0, 400_000, n)).astype(np.uint32)
    ms_to_idx = np.searchsorted(t, np.arange(0, 401) * 1000, side="left").astype(np.int64)
    h5 = {"events/x": rng.integers(0, W, n).astype(np.uint16), "events/y": rng.integers(0, H, n).astype(np.uint16),
          "events/t": t, "events/p": rng.integers(0, 2, n).astype(np.uint8), "ms_to_idx": ms_to_idx,
          "t_offset": np.array(1_000_000, dtype=np.int64)}
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rmap = (np.stack([xx, yy], -1) + rng.uniform(-0.75, 0.75, (H, W, 2))).astype(np.float32)
    seq = object.__new__(Sequence)
    seq.mode, seq.height, seq.width, seq.resize, seq.shape_resize = "train", H, W, False, None
    seq.nr_events_data, seq.num_bins, seq.nr_events_per_data = nd, 5, per
    seq.event_representation, seq.separate_pol, seq.normalize_event = "voxel_grid", False, False
    seq.voxel_grid = VoxelGrid(5, H, W, normalize=False)
    seq.locations, seq.semseg_num_classes, seq.augmentation, seq.fixed_duration = ["left"], 11, True, False
    seq.config_option, seq.pl_sources, seq.superpixel_sources, seq.if_sam_distillation = "frame2voxel", "pl_fcclip_rgb", "sp_sam_rgb", False
    seq.rectify_ev_maps = {"left": rmap}
    seq.event_slicers = {"left": EventSlicer(h5)}
    seq.timestamps = np.array([1_150_000, 1_250_000, 1_390_000], dtype=np.int64)
    base = Path(TMP) / "seq0"
    paths = []
    for i in range(3):
        name = "%06d.png" % i
        for sub, mode, hi in (("semantic/left/11classes", "L", 11), ("pl_fcclip_rgb/left", "L", 11), ("sp_sam_rgb/left", "L", 100)):
            d = base / sub
            d.mkdir(parents=True, exist_ok=True)
            Image.fromarray(rng.integers(0, hi, (440, 640)).astype(np.uint8), mode).save(d / name)
        d = base / "images_aligned/left"
        d.mkdir(parents=True, exist_ok=True)
        Image.fromarray(rng.integers(0, 255, (440, 640, 3)).astype(np.uint8), "RGB").save(d / name)
        paths.append(str(base / "semantic/left/11classes" / name))
    seq.label_pathstrings = paths
    return seq

if sys.argv[3] == "reference":                      # the unpatched reference, one thread (serial put_, SURVEY.md 0.5)
    torch.set_num_threads(1)
    sys.path.insert(0, "/root/reference")
    ds = types.ModuleType("datasets"); ds.__path__ = ["/root/reference/datasets"]; sys.modules["datasets"] = ds
    from DSEC.dataset.sequence_ov import Sequence
    from DSEC.dataset.representations import VoxelGrid
    from openess_b200.DSEC.utils.eventslicer import EventSlicer     # dict-backed slicer (golden-tested against the reference's)
    seq = build_sequence(Sequence, EventSlicer, VoxelGrid, np.random.default_rng(3))
    out = {}
    for i in range(3):
        random.seed(100 + i); torch.manual_seed(100 + i)
        item = seq[i]
        out["ev%d" % i] = item[0].numpy()
        out["label%d" % i], out["frame%d" % i], out["pl%d" % i], out["sp%d" % i] = (a.numpy() for a in item[1:5])
    np.savez(os.path.join(TMP, "reference.npz"), **out)
else:                                               # patched: stock DataLoader with one forked worker
    from openess_b200.patch import patch_reference
    done = patch_reference("/root/reference")
    assert not isinstance(done["DSEC.dataset.sequence_ov.Sequence.__getitem__"], Exception), done["DSEC.dataset.sequence_ov.Sequence.__getitem__"]
    k = "training.pretrain_trainer.OpenESSPretrainModel.task_train_step"
    assert not isinstance(done[k], Exception), done[k]
    import training.pretrain_trainer as pt          # the reference's trainer module: methods rebound in place
    assert hasattr(pt.OpenESSPretrainModel.task_train_step, "__wrapped__") and hasattr(pt.OpenESSPretrainModel.train_step, "__wrapped__")
    assert pt.OpenESSPretrainModel.task_train_step.__wrapped__.__code__.co_filename.startswith("/root/reference")
    assert pt.SemSegE2VID.__module__.startswith("openess_b200") and pt.ImageReconstructor.__module__.startswith("openess_b200")
    import DSEC.dataset.sequence_ov as so
    from DSEC.dataset.representations import VoxelGrid
    from DSEC.utils.eventslicer import EventSlicer
    from openess_b200.training.pretrain_step import RawEvents
    seq = build_sequence(so.Sequence, EventSlicer, VoxelGrid, np.random.default_rng(3))

    class Seeded(torch.utils.data.Dataset):         # the same augmentation draws as the reference run
        def __len__(self): return 3
        def __getitem__(self, i):
            random.seed(100 + i); torch.manual_seed(100 + i)
            return seq[i]
    loader = torch.utils.data.DataLoader(Seeded(), batch_size=3, num_workers=1, pin_memory=False, shuffle=False, drop_last=True)
    batch = next(iter(loader))
    ev = batch[0]
    assert isinstance(ev, RawEvents), type(ev)
    assert not torch.cuda.is_initialized()
    np.savez(os.path.join(TMP, "patched.npz"), x=ev.x.numpy(), y=ev.y.numpy(), t=ev.t.numpy(), p=ev.p.numpy(),
             fo=ev.frame_offsets.numpy(), rmap=ev.rectify_map.numpy(), crop_h=np.array(ev.crop_h),
             flip=np.zeros(3, np.uint8) if ev.flip is None else ev.flip.numpy(),
             label=batch[1].numpy(), frame=batch[2].numpy(), pl=batch[3].numpy(), sp=batch[4].numpy())
print("OK")
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_sequence_getitem_returns_raw_slabs_from_forked_worker(tmp_path):
    from oracle import oracle as orc
    orc.build()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER_CODE)
    for which in ("reference", "patched"):
        r = subprocess.run([sys.executable, str(script), ROOT, str(tmp_path), which], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "OK" in r.stdout, (which, r.stderr[-3000:])
    ref = np.load(tmp_path / "reference.npz")
    got = np.load(tmp_path / "patched.npz")
    fo, crop_h = got["fo"], int(got["crop_h"])
    assert len(fo) == 3 * 4 + 1 and got["x"].dtype == np.uint16 and got["t"].dtype == np.int64 and got["p"].dtype == np.uint8
    flips = []
    for b in range(3):
        frames = []
        for j in range(4):
            s, e = int(fo[4 * b + j]), int(fo[4 * b + j + 1])
            xo, yo, po, to = orc.dsec_rectify_tnorm(got["x"][s:e], got["y"][s:e], got["t"][s:e], got["p"][s:e], got["rmap"])
            frames.append(orc.voxel_trilinear(xo, yo, po, to, 5, 480, 640))
        dense = np.concatenate(frames, 0)[:, :crop_h, :]
        if got["flip"][b]:
            dense = dense[:, :, ::-1]
        flips.append(int(got["flip"][b]))
        assert dense.tobytes() == np.ascontiguousarray(ref[f"ev{b}"]).tobytes(), f"sample {b}: slab -> voxel grid differs from the reference's dense tensor"
        for k, idx in (("label", 1), ("frame", 2), ("pl", 3), ("sp", 4)):
            assert np.array_equal(got[k][b], ref[f"{k}{b}"]), (k, b)        # the rest of __getitem__ ran unmodified
    assert any(flips), "the seeds are chosen so that the augmentation flips at least one sample"


class _Trainer:
    """Shaped like OpenESSPretrainModel after init (training/pretrain_trainer.py:81-105, 211-274): what drop_in reads."""

    def task_train_step(self, batch):
        raise AssertionError("the reference formulation must not run for frame2voxel on CUDA")

    def train_step(self, input_batch):                       # pretrain_trainer.py:324-361, non-AMP branch
        for key in ("optimizer_voxel", "optimizer_frame"):
            self.optimizers_dict[key].zero_grad()
        final, losses, outputs = self.task_train_step(input_batch)
        final.backward()
        for key in ("optimizer_voxel", "optimizer_frame"):
            self.optimizers_dict[key].step()
        return losses, outputs, final


@pytest.mark.gpu
def test_rebound_trainer_methods_run_the_fused_step_and_match_golden():
    from types import SimpleNamespace
    from conftest import load_golden
    from test_pretrain_step import _models
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.models import image_model as im
    from openess_b200.models import style_networks as sn
    from openess_b200.training import drop_in
    from openess_b200.utils.loss_functions import NCELoss, TaskLoss
    z = load_golden("pretrain_step")
    dev = torch.device("cuda:0")
    e2vid, back, teacher, opts, K = _models(dev)
    S, steps = int(z["S"]), int(z["steps"])
    event, frame, pl, sp = (torch.from_numpy(z[k]) for k in ("event", "frame", "pl", "sp"))     # CPU, as a DataLoader hands them over
    H, W = event.shape[-2:]

    class T(_Trainer):
        pass
    T.task_train_step = drop_in.wrap_task_train_step(T.task_train_step)
    T.train_step = drop_in.wrap_train_step(T.train_step)
    tr = T()
    tr.device = dev
    tr.epoch_count = 0
    tr.settings = SimpleNamespace(config_option="frame2voxel", use_amp=False, unfrozen_e2vid=False, if_switchable_train=False,
                                  nr_events_data_b=steps, input_channels_b=5, superpixel_size=S, weight_task_loss=1.0,
                                  if_spatial_contrastive=True, if_dense_clip_supervision=True)
    tr.models_dict = {"front_sensor_b": e2vid, "back_end": back, "model_frame": teacher}
    tr.reconstructor = ImageReconstructor(e2vid, H, W, 5, dev, opts)
    tr.task_loss = TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    tr.nce_loss = NCELoss(temperature=0.07)
    tr.optimizers_dict = {"optimizer_voxel": torch.optim.AdamW([p for p in back.parameters() if p.requires_grad], lr=1e-3),
                          "optimizer_frame": torch.optim.AdamW([p for p in teacher.parameters() if p.requires_grad], lr=1e-3)}
    im.USE_TENSOR_CORES = False
    sn.TRAIN_ON_TENSOR_CORES = False
    try:
        w0 = back.decoder_ch256[0].weight.detach().clone()
        losses, outputs, final = tr.train_step((event, None, frame, pl, sp))
        assert float(final) == pytest.approx(float(z["step_total"]), rel=3e-4)                   # == the reference trainer's own step
        assert float(losses["contrastive_nce_loss"]) == pytest.approx(float(z["nce"]), rel=3e-4)
        assert tr._oess_step.optimizers_dict is tr.optimizers_dict                              # the trainer's own optimisers stepped
        assert not torch.equal(back.decoder_ch256[0].weight, w0)
        ref_after = z["after__back_end.decoder_ch256.0.weight"]
        d = np.abs(back.decoder_ch256[0].weight.detach().cpu().numpy() - ref_after)
        assert float(d.max()) <= 2.1e-3 and float((d < 2e-5).mean()) > 0.98
        # other configurations fall through to the reference's method
        tr.settings.config_option = "frame2recon"
        with pytest.raises(AssertionError, match="reference formulation"):
            tr.task_train_step((event, None, frame, pl, sp))
    finally:
        im.USE_TENSOR_CORES = True
        sn.TRAIN_ON_TENSOR_CORES = True
