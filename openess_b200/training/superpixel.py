"""The superpixel-pooled contrastive block that the reference inlines six times in its trainers
(training/pretrain_trainer.py:397-417, 445-465, 487-506; training/openess_trainer.py:406-429, 505-529),
as one fused segment-reduce per feature map instead of a sparse one-hot matmul on permuted copies."""
import torch

from ..losses import superpixel_pool


def pooled_pair(feat_a, feat_b, superpixels, superpixel_size):
    """-> (k, q): superpixel mean-pooled features of both maps, [M, C] each; M = max id' + 1 like the reference."""
    B = feat_a.shape[0]
    off = torch.arange(0, B * superpixel_size, superpixel_size, device=superpixels.device)[:, None, None]
    M = int((superpixels + off).max().item()) + 1          # torch.sparse_coo_tensor infers the same size
    k = superpixel_pool(feat_a, superpixels, superpixel_size, M)
    q = superpixel_pool(feat_b, superpixels, superpixel_size, M)
    return k, q


def spatial_contrastive_loss(feat_voxel, feat_other, superpixels, superpixel_size, nce_loss):
    """pretrain_trainer.py:445-467: loss_contrastive_nce = nce_loss(k, q)."""
    k, q = pooled_pair(feat_voxel, feat_other, superpixels, superpixel_size)
    return nce_loss(k, q)
