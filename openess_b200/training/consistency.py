"""The two consistency losses of the reference's OpenESSModel trainer (training/openess_trainer.py:456-462, also
:398-404 and :497-503) as fused kernels: mean |a - b| over the 256-channel feature maps and
mean(1 - cosine_similarity(logits_a, logits_b, dim=1))."""
import torch

from .. import losses as _ops


class L1Loss(torch.nn.Module):
    """Drop-in for `torch.nn.L1Loss()` as constructed at openess_trainer.py:92 (mean reduction)."""

    def forward(self, a, b):
        return _ops.l1_mean(a, b)


def prediction_consistency(logits_a, logits_b):
    """torch.mean(1 - f.cosine_similarity(logits_a, logits_b, dim=1))  (openess_trainer.py:460)."""
    return _ops.cosine_consistency(logits_a, logits_b)
