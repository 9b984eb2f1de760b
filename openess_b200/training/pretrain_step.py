"""The OpenESS stage-1 pretraining step (config_option 'frame2voxel': frame-to-event InfoNCE on superpixels +
text-to-event Dice/CE against FC-CLIP pseudo-labels), SURVEY.md 8a row a19.

Mirror of training/pretrain_trainer.py: `createOptimizerDict` :211-243 (two AdamW), `train_step` :324-361,
`task_train_step` :364-372 + :427-472 (frame2voxel branch), `trainTaskStepPretrain` :550-562 -- same attribute names
(`models_dict`, `optimizers_dict`, `reconstructor`, `task_loss`, `nce_loss`), same batch tuple
`(event, label, frame, pl, superpixels, ...)`, same returned `(losses, outputs, final_loss)`.

What changes on the B200:
  * batch[0] may be the reference's dense event tensor [B, 20*5, H, W] OR a `RawEvents` slab (the raw DSEC records of
    the sample windows): the slab is rectified, time-normalised and voxelised on the device in one batched call
    (F = B * 20 frames), replacing the 123 MB / sample host->device copy by 18 MB / sample;
  * E2VID (20 recurrent steps) and the frozen teacher run on the tcgen05 kernels, sync-free;
  * `SemSegE2VID.forward_pooled` + `superpixel_pool` replace the sparse one-hot matmuls on permuted 2.3 GB copies
    (:446-465) -- k and q are produced by fused segment reductions, the 256 / 512-channel event-branch maps never exist;
  * `NCELoss` / `TaskLoss` are the fused kernels; under data parallelism `parallel.GradientReducer` averages the gradients
    of the two trainable modules over NCCL in buckets that are all-reduced WHILE backward is still running (the
    reference is single-GPU, README.md:303).
"""
import os
from collections import namedtuple

import torch

from .. import parallel as _parallel
from .. import voxel as _voxel
from ..losses import superpixel_pool

# raw DSEC records of B samples x nr_events_data windows: x, y uint16; t uint32 or int64 (us); p uint8; frame_offsets
# int64 [B * nr_events_data + 1]; rectify_map float32 [Hs, Ws, 2] (sequence_ov.py:204-210); sensor (Hs, Ws); crop rows;
# flip: None or uint8 / bool [B], samples whose event tensor the loader's augmentation flipped horizontally (:387-389)
RawEvents = namedtuple("RawEvents", "x y t p frame_offsets rectify_map sensor_hw crop_h flip", defaults=(None,))


def assemble_event_tensor(ev, device, nr_events_data=20, C=5, out=None):
    """batch[0] of a trainer step -> dense [B, nr_events_data * C, crop_h, W] on `device`: a `RawEvents` slab is rectified,
    time-normalised and voxelised there in one batched call (F = B * nr_events_data frames; sequence_ov.py:282-307); the
    reference's dense tensor is just moved."""
    if not isinstance(ev, RawEvents):
        return ev.to(device)                                   # reference format: dense [B, 20*5, H, W]
    Hs, Ws = ev.sensor_hw
    x, y, t, p = (a.to(device, non_blocking=True) for a in (ev.x, ev.y, ev.t, ev.p))
    fo = ev.frame_offsets.to(device, non_blocking=True)
    F = fo.numel() - 1
    if out is not None and (tuple(out.shape) != (F, C, Hs, Ws) or out.device != x.device):
        out = None                                             # `out`: a persistent [F, C, Hs, Ws] buffer (stable address for a CUDA graph)
    grids = _voxel.dsec_events_to_voxel_grid(x, y, t, p, ev.rectify_map.to(device), C, frame_offsets=fo,
                                             mode="ordered", out=out)               # [F, C, Hs, Ws]
    B = F // nr_events_data
    dense = grids.view(B, nr_events_data * C, Hs, Ws)
    if ev.flip is not None:                                    # sequence_ov.py:387-389 torch.flip(event_tensor, [2]) per sample
        from ..DSEC.dataset.augment import hflip_rows_
        hflip_rows_(dense, torch.as_tensor(ev.flip).to(device).to(torch.uint8))
    return dense[:, :, :ev.crop_h, :]                          # sequence_ov.py:307 bottom crop, a view


class OpenESSPretrainStep:
    def __init__(self, reconstructor, back_end, model_frame, task_loss, nce_loss, *, nr_events_data_b=20,
                 input_channels_b=5, superpixel_size=100, weight_task_loss=1.0, if_spatial_contrastive=True,
                 if_dense_clip_supervision=True, lr_voxel=5e-4, lr_frame=5e-4, device=None, data_parallel=False,
                 optimizers_dict=None):
        self.reconstructor = reconstructor
        self.models_dict = {"front_sensor_b": reconstructor.model, "back_end": back_end, "model_frame": model_frame}
        self.task_loss, self.nce_loss = task_loss, nce_loss
        self.nr_events_data_b, self.input_channels_b = nr_events_data_b, input_channels_b
        self.superpixel_size, self.weight_task_loss = superpixel_size, weight_task_loss
        self.if_spatial_contrastive, self.if_dense_clip_supervision = if_spatial_contrastive, if_dense_clip_supervision
        self.device = device if device is not None else next(back_end.parameters()).device
        self.data_parallel = data_parallel
        for p in reconstructor.model.parameters():             # pretrain_trainer.py:154-157: frozen E2VID
            p.requires_grad = False
        # :231-243
        params_voxel = [p for p in back_end.parameters() if p.requires_grad]
        params_frame = [p for p in model_frame.parameters() if p.requires_grad]
        fused = self.device.type == "cuda"
        if optimizers_dict is not None:                        # a reference trainer's own optimisers (drop_in.py)
            self.optimizers_dict = optimizers_dict
        else:
            self.optimizers_dict = {"optimizer_voxel": torch.optim.AdamW(params_voxel, lr=lr_voxel, fused=fused),
                                    "optimizer_frame": torch.optim.AdamW(params_frame, lr=lr_frame, fused=fused)}
        self._trainable = params_voxel + params_frame
        self._reducer = _parallel.GradientReducer(self._trainable) if data_parallel else None
        # the frozen recurrent encoder loop as one CUDA graph (OESS_E2VID_GRAPH=0: eager); only when the encoder is frozen
        self._grids = None
        self._graph_loop = None
        e2vid = reconstructor.model
        frozen = not any(p.requires_grad for p in e2vid.parameters())
        if os.environ.get("OESS_E2VID_GRAPH", "1") != "0" and self.device.type == "cuda" and frozen:
            from .graphs import GraphedEncoderLoop
            self._graph_loop = GraphedEncoderLoop(reconstructor, nr_events_data_b, input_channels_b)

    # ---- sample assembly on the device (replaces Sequence.__getitem__'s voxel branch, sequence_ov.py:282-307) ----
    def event_tensor(self, ev):
        if not (getattr(self, "_graph_loop", None) is not None and isinstance(ev, RawEvents)):
            return assemble_event_tensor(ev, self.device, self.nr_events_data_b, self.input_channels_b)
        # CUDA-graph mode: the voxeliser writes into a persistent buffer, so the captured encoder loop reads a stable address
        F = ev.frame_offsets.numel() - 1
        shape = (F, self.input_channels_b) + tuple(ev.sensor_hw)
        if self._grids is None or tuple(self._grids.shape) != shape:
            self._grids = torch.empty(shape, dtype=torch.float32, device=self.device)
        return assemble_event_tensor(ev, self.device, self.nr_events_data_b, self.input_channels_b, out=self._grids)

    # ---- pretrain_trainer.py:550-562 ----
    def trainTaskStepPretrain(self, content_features, pl, superpixels, losses):
        content_features = {k: v.detach() for k, v in content_features.items()}
        back_end = self.models_dict["back_end"]
        if self.if_spatial_contrastive and hasattr(back_end, "forward_pooled"):
            pred, k = back_end.forward_pooled(content_features, superpixels, self.superpixel_size)
        else:
            pred, k = back_end(content_features)[0], None
        loss = self.task_loss(pred[1], pl) * self.weight_task_loss
        losses["dense_clip_loss"] = loss.detach()
        return loss, pred, k

    # ---- pretrain_trainer.py:364-372, 427-472 ----
    def task_train_step(self, batch):
        losses, outputs, t_loss = {}, {}, 0.0
        for name, m in self.models_dict.items():
            m.train()                                          # :370-371 (the "frozen" teacher's BN uses batch statistics)
            if name in ("front_sensor_b", "model_clip"):
                m.eval()                                       # :372-377
        event = self.event_tensor(batch[0])
        frame = batch[2].to(self.device)
        pl = batch[3].to(self.device)
        superpixels = batch[4].to(self.device) if self.if_spatial_contrastive else None
        model_frame = self.models_dict["model_frame"]
        fused_q = self.if_spatial_contrastive and frame.is_cuda and hasattr(model_frame, "forward_pooled")
        feat_frame = None if fused_q else model_frame(frame)                                 # :434
        C = self.input_channels_b
        if getattr(self, "_graph_loop", None) is not None and event.is_cuda:
            latent_real = self._graph_loop(event)              # the frozen 20-step encoder loop as one CUDA graph (training/graphs.py)
        else:
            self.reconstructor.last_states_for_each_channel = {"grayscale": None}
            for i in range(self.nr_events_data_b):
                _, _, latent_real = self.reconstructor.update_reconstruction(event[:, i * C:(i + 1) * C])
        loss_dense, pred, k = self.trainTaskStepPretrain(latent_real, pl, superpixels, losses)
        if self.if_spatial_contrastive:
            M = k.shape[0]
            if fused_q:                                                                      # :434 + :461-463 in one pass
                q = model_frame.forward_pooled(frame, superpixels, self.superpixel_size, M)
            else:
                q = superpixel_pool(feat_frame, superpixels, self.superpixel_size, M)       # :461-463
            loss_nce = self.nce_loss(k, q)                                                  # :467
            losses["contrastive_nce_loss"] = loss_nce.detach()
            t_loss = t_loss + loss_nce
        if self.if_dense_clip_supervision:
            t_loss = t_loss + loss_dense
        outputs["pred"] = pred
        return t_loss, losses, outputs

    # ---- pretrain_trainer.py:324-361 ----
    def train_step(self, input_batch):
        for opt in self.optimizers_dict.values():
            opt.zero_grad(set_to_none=True)
        final_loss, losses, outputs = self.task_train_step(input_batch)
        if self._reducer is not None:
            self._reducer.prepare()
        final_loss.backward()
        if self._reducer is not None:
            self._reducer.finish()                         # bucketed all-reduce, overlapped with backward from step 2 on
        for opt in self.optimizers_dict.values():
            opt.step()
        return losses, outputs, final_loss
