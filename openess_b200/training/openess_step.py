"""The step of the reference's `OpenESSModel` trainer (training/openess_trainer.py), branch `config_option == 'frame2recon'`:
two DeepLabv3-ResNet-50 networks (frame and E2VID reconstruction) supervised by the FC-CLIP pseudo-labels, tied by an L1 feature
consistency, a cosine prediction consistency and the superpixel InfoNCE (:476-529), two AdamW optimisers (:262-272, :339-358).

Only this branch is built: the reference's `recon2voxel` branch raises NameError (`superpixel` vs `superpixels`, :379 / :408-409)
and its `frame2voxel` branch never sets the `contrastive_nce_loss` that `trainEpoch` reads (:464-475 vs :307-308), SURVEY.md
Appendix B.14.  Same attribute names (`models_dict`, `optimizers_dict`, `task_loss`, `l1_loss`, `nce_loss`), same batch tuple
`(frame, label, recon, pl, superpixels, ...)`, same returned `(losses, outputs, final_loss)` and loss-dict keys.

On the B200 the loss side runs on the fused kernels (Dice + CE in one pass per network, `oess_l1_mean`, `oess_cos_consistency`,
segment-reduce pooling, InfoNCE); the two trainable ResNet-50 DeepLabs run the module mirrors' autograd graph."""
import torch

from .. import parallel as _parallel
from ..losses import superpixel_pool
from .consistency import L1Loss, prediction_consistency


class OpenESSStep:
    SUPERPIXEL_SIZE = 30                                        # hard-coded in the reference (:505)

    def __init__(self, model_frame, model_recon, task_loss, nce_loss, *, weight_task_loss=1.0, if_spatial_contrastive=True,
                 lr_recon=5e-4, lr_frame=5e-4, device=None, data_parallel=False, optimizers_dict=None):
        self.models_dict = {"model_recon": model_recon, "model_frame": model_frame}
        self.task_loss, self.nce_loss, self.l1_loss = task_loss, nce_loss, L1Loss()
        self.weight_task_loss, self.if_spatial_contrastive = weight_task_loss, if_spatial_contrastive
        self.device = device if device is not None else next(model_frame.parameters()).device
        params_recon = [p for p in model_recon.parameters() if p.requires_grad]          # :263-267
        params_frame = [p for p in model_frame.parameters() if p.requires_grad]
        fused = self.device.type == "cuda"
        self.optimizers_dict = optimizers_dict if optimizers_dict is not None else {
            "optimizer_recon": torch.optim.AdamW(params_recon, lr=lr_recon, fused=fused),
            "optimizer_frame": torch.optim.AdamW(params_frame, lr=lr_frame, fused=fused)}
        self._reducer = _parallel.GradientReducer(params_recon + params_frame) if data_parallel else None

    # ---- openess_trainer.py:360-372, 476-529 ----
    def task_train_step(self, batch):
        losses, outputs, t_loss = {}, {}, 0.0
        for m in self.models_dict.values():
            m.train()
        frame = batch[0].to(self.device)
        recon = batch[2].to(self.device)
        pl = batch[3].to(self.device)
        superpixels = batch[4].to(self.device) if self.if_spatial_contrastive else None
        logits_frame, feat_frame = self.models_dict["model_frame"](frame)                # :483
        loss_pred_frame = self.task_loss(logits_frame, pl) * self.weight_task_loss
        losses["semseg_frame_loss"] = loss_pred_frame.detach()
        t_loss = t_loss + loss_pred_frame
        logits_recon, feat_recon = self.models_dict["model_recon"](recon)                # :489
        loss_pred_recon = self.task_loss(logits_recon, pl) * self.weight_task_loss
        losses["semseg_recon_loss"] = loss_pred_recon.detach()
        t_loss = t_loss + loss_pred_recon
        loss_cons_feat = self.l1_loss(feat_frame, feat_recon)                            # :495
        losses["cons_feat_loss"] = loss_cons_feat.detach()
        t_loss = t_loss + loss_cons_feat
        loss_cons_pred = prediction_consistency(logits_frame, logits_recon)              # :499
        losses["cons_pred_loss"] = loss_cons_pred.detach()
        t_loss = t_loss + loss_cons_pred
        if self.if_spatial_contrastive:                                                  # :503-529
            k = superpixel_pool(feat_recon, superpixels, self.SUPERPIXEL_SIZE)
            q = superpixel_pool(feat_frame, superpixels, self.SUPERPIXEL_SIZE, k.shape[0])
            loss_nce = self.nce_loss(k, q)
            losses["contrastive_nce_loss"] = loss_nce.detach()
            t_loss = t_loss + loss_nce
        return t_loss, losses, outputs

    # ---- openess_trainer.py:339-358 ----
    def train_step(self, input_batch):
        for key in ("optimizer_recon", "optimizer_frame"):
            self.optimizers_dict[key].zero_grad(set_to_none=True)
        final_loss, losses, outputs = self.task_train_step(input_batch)
        if self._reducer is not None:
            self._reducer.prepare()
        final_loss.backward()
        if self._reducer is not None:
            self._reducer.finish()
        for key in ("optimizer_recon", "optimizer_frame"):
            self.optimizers_dict[key].step()
        return losses, outputs, final_loss
