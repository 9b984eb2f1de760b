"""Zero-edit drop-in of the fused pretraining step and of device-side sample assembly.

`install(pretrain_trainer_module, sequence_module=None)` -- called by `patch_reference` -- rebinds, on the REFERENCE's own classes:

  * `OpenESSPretrainModel.task_train_step` / `.train_step` (training/pretrain_trainer.py:324-361, 364-372, 427-472, 550-562):
    for `config_option == 'frame2voxel'` (without AMP / switchable pseudo-labels) the step runs through
    `OpenESSPretrainStep`, built lazily from the trainer's OWN modules, losses, settings and optimisers; every other
    configuration falls through to the reference's method.  Under torchrun the gradients of the two trainable modules are
    all-reduced in buckets overlapped with backward (parallel.GradientReducer).
  * `Sequence.__getitem__` (DSEC/dataset/sequence_ov.py:225-463): for the count-windowed voxel-grid configurations (every
    shipped YAML: `fixed_duration: False`, `event_representation: 'voxel_grid'`) the worker no longer voxelises -- CUDA cannot
    be initialised in a forked DataLoader worker, and the dense [100, 440, 640] tensor is 123 MB per sample.  The reference's
    `__getitem__` still runs UNMODIFIED (label / frame / pseudo-label / superpixel loading, the `random` draws of the
    augmentation), but with `rectify_events` and `generate_event_tensor` of that instance replaced by recorders, so the event
    tensor it returns is an untouched zero placeholder; the wrapper swaps it for a `RawSlab` (the raw records of the 20 windows,
    18 MB) and notes whether the augmentation flipped it.  `collate_raw_slabs` (registered in torch's `default_collate_fn_map`, so the
    reference's stock `DataLoader(...)` call needs no `collate_fn`) concatenates the slabs of a batch into one `RawEvents`,
    which the rebound `task_train_step` voxelises on the device in the MAIN process.
"""
import threading

import numpy as np
import torch

from .. import parallel as _parallel
from .pretrain_step import OpenESSPretrainStep, RawEvents, assemble_event_tensor

_MARK = 7.0


class RawSlab:
    """Raw DSEC records of ONE sample (nr_events_data windows): what a DataLoader worker returns instead of the voxel grids."""
    __slots__ = ("x", "y", "t", "p", "offsets", "rectify_map", "sensor_hw", "crop_h", "flip")

    def __init__(self, x, y, t, p, offsets, rectify_map, sensor_hw, crop_h, flip):
        self.x, self.y, self.t, self.p, self.offsets = x, y, t, p, offsets
        self.rectify_map, self.sensor_hw, self.crop_h, self.flip = rectify_map, sensor_hw, crop_h, flip

    def __getstate__(self):
        return tuple(getattr(self, k) for k in self.__slots__)

    def __setstate__(self, st):
        for k, v in zip(self.__slots__, st):
            setattr(self, k, v)


def collate_raw_slabs(batch, *, collate_fn_map=None):
    """B RawSlab -> one RawEvents (CPU tensors; the trainer's `.to(self.device)` is replaced by the device assembly)."""
    offs, base = [0], 0
    for s in batch:
        o = np.asarray(s.offsets, dtype=np.int64)
        offs.extend((o[1:] + base).tolist())
        base += int(o[-1])
    cat = lambda k, dt: torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(getattr(s, k)) for s in batch]).astype(dt, copy=False)))  # noqa: E731
    flip = torch.tensor([bool(s.flip) for s in batch], dtype=torch.uint8)
    return RawEvents(cat("x", np.uint16), cat("y", np.uint16), cat("t", np.int64), cat("p", np.uint8),
                     torch.tensor(offs, dtype=torch.int64), torch.from_numpy(np.asarray(batch[0].rectify_map, dtype=np.float32)),
                     tuple(batch[0].sensor_hw), int(batch[0].crop_h), flip if bool(flip.any()) else None)


def register_collate():
    from torch.utils.data._utils import collate as _c
    _c.default_collate_fn_map[RawSlab] = collate_raw_slabs


# ------------------------------------------------------------------------------------------ Sequence.__getitem__
_tls = threading.local()


def _raw_mode_ok(seq):
    return (getattr(seq, "config_option", None) in ("frame2voxel", "recon2voxel") and not getattr(seq, "fixed_duration", True)
            and getattr(seq, "event_representation", "voxel_grid") == "voxel_grid" and not getattr(seq, "resize", False)
            and list(getattr(seq, "locations", ["left"])) == ["left"])


def wrap_sequence_getitem(orig_getitem):
    def __getitem__(self, index):
        if not _raw_mode_ok(self):
            return orig_getitem(self, index)
        rec = {}

        def rectify_events(x, y, location):                       # sequence_ov.py:204-210: record, do not gather
            rec["x"], rec["y"] = np.asarray(x), np.asarray(y)
            return np.zeros((x.shape[0], 2), dtype=np.float32)

        def generate_event_tensor(job_id, events, event_tensor, nr_events_per_data):      # :212-223: record, do not voxelise
            if job_id == 0:
                rec["t"] = np.asarray(events[:, 2])
                rec["p"] = np.asarray(events[:, 3])
                rec["per"] = int(nr_events_per_data)
                event_tensor[0, 0, 0] = _MARK                     # finds out whether the augmentation flips the tensor

        self.rectify_events, self.generate_event_tensor = rectify_events, generate_event_tensor
        try:
            out = orig_getitem(self, index)
        finally:
            del self.rectify_events, self.generate_event_tensor   # back to the class attributes
        ev = out[0]
        flipped = bool(ev[0, 0, -1] == _MARK) and not bool(ev[0, 0, 0] == _MARK)
        nd, per = int(self.nr_events_data), rec["per"]
        n = nd * per                                              # generate_event_tensor reads events[j * per:(j + 1) * per]
        slab = RawSlab(rec["x"][:n].astype(np.uint16, copy=False), rec["y"][:n].astype(np.uint16, copy=False),
                       rec["t"][:n].astype(np.int64), rec["p"][:n].astype(np.uint8),
                       np.arange(0, n + 1, per, dtype=np.int64), self.rectify_ev_maps["left"],
                       (int(self.height), int(self.width)), int(ev.shape[1]), flipped)
        return (slab,) + tuple(out[1:])
    __getitem__.__wrapped__ = orig_getitem
    return __getitem__


# ------------------------------------------------------------------------------------------ trainer methods
def _fused_ok(tr):
    s = tr.settings
    return (getattr(s, "config_option", None) == "frame2voxel" and not getattr(s, "use_amp", False)
            and not getattr(s, "unfrozen_e2vid", False)
            and not (getattr(s, "if_switchable_train", False) and getattr(tr, "epoch_count", 0) >= 5)
            and next(tr.models_dict["back_end"].parameters()).is_cuda)


def _fused_step(tr):
    st = getattr(tr, "_oess_step", None)
    if st is None:
        s = tr.settings
        st = OpenESSPretrainStep(tr.reconstructor, tr.models_dict["back_end"], tr.models_dict["model_frame"], tr.task_loss,
                                 tr.nce_loss, nr_events_data_b=s.nr_events_data_b, input_channels_b=s.input_channels_b,
                                 superpixel_size=s.superpixel_size, weight_task_loss=s.weight_task_loss,
                                 if_spatial_contrastive=s.if_spatial_contrastive,
                                 if_dense_clip_supervision=s.if_dense_clip_supervision, device=tr.device,
                                 data_parallel=_parallel.world_size() > 1, optimizers_dict=tr.optimizers_dict)
        st.models_dict = tr.models_dict                           # the trainer's own dict (model_clip etc. stay where they are)
        tr._oess_step = st
    return st


def wrap_task_train_step(orig):
    def task_train_step(self, batch):
        if not _fused_ok(self):
            if isinstance(batch[0], RawEvents):                   # the reference path wants the dense tensor
                s = self.settings
                batch = (assemble_event_tensor(batch[0], self.device, s.nr_events_data_b, s.input_channels_b),) + tuple(batch[1:])
            return orig(self, batch)
        return _fused_step(self).task_train_step(batch)
    task_train_step.__wrapped__ = orig
    return task_train_step


def wrap_train_step(orig):
    def train_step(self, input_batch):
        if not _fused_ok(self) or _parallel.world_size() == 1:
            return orig(self, input_batch)                        # :324-361 as written (it calls the rebound task_train_step)
        st = _fused_step(self)
        for key in ("optimizer_voxel", "optimizer_frame"):        # :340-342
            self.optimizers_dict[key].zero_grad()
        final_loss, losses, outputs = self.task_train_step(input_batch)
        st._reducer.prepare()
        final_loss.backward()
        st._reducer.finish()
        for key in ("optimizer_voxel", "optimizer_frame"):        # :355-359
            self.optimizers_dict[key].step()
        return losses, outputs, final_loss
    train_step.__wrapped__ = orig
    return train_step


def install(pretrain_trainer_module=None, sequence_module=None):
    done = {}
    if pretrain_trainer_module is not None:
        T = pretrain_trainer_module.OpenESSPretrainModel
        if not hasattr(T.task_train_step, "__wrapped__"):
            T.task_train_step = wrap_task_train_step(T.task_train_step)
            T.train_step = wrap_train_step(T.train_step)
        done["training.pretrain_trainer.OpenESSPretrainModel.task_train_step"] = T.task_train_step
        done["training.pretrain_trainer.OpenESSPretrainModel.train_step"] = T.train_step
    if sequence_module is not None:
        S = sequence_module.Sequence
        if not hasattr(S.__getitem__, "__wrapped__"):
            S.__getitem__ = wrap_sequence_getitem(S.__getitem__)
        register_collate()
        done["DSEC.dataset.sequence_ov.Sequence.__getitem__"] = S.__getitem__
    return done
