"""Drop-in mirror of the reference's utils/loss_functions.py (same class names, constructor arguments and
call signatures), computing on the GPU through libopeness_b200 with fused forward / backward kernels."""
import torch

from .. import losses as _losses


class TaskLoss(torch.nn.Module):
    """loss_functions.py:6-24: Dice (if 'dice' in losses) + CrossEntropy(ignore_index) (if 'cross_entropy')."""

    def __init__(self, losses=['cross_entropy'], gamma=2.0, num_classes=13, alpha=None, weight=None,
                 ignore_index=None, reduction='mean'):
        super(TaskLoss, self).__init__()
        self.losses = losses
        self.weight = weight
        self.gamma = gamma
        self.alpha = alpha
        self.ignore_index = ignore_index
        self.num_classes = num_classes
        self.dice_loss = DiceLoss(num_classes=num_classes, ignore_index=self.ignore_index)
        # exact global-batch semantics under data parallelism: set to a callable that all-reduces the
        # float64 partial sums in place (openess_b200.parallel.allreduce_sum_); None = local batch
        self.reduce_partials = None

    def forward(self, predict, target):
        w_dice = 1.0 if 'dice' in self.losses else 0.0
        w_ce = 1.0 if 'cross_entropy' in self.losses else 0.0
        if w_dice == 0.0 and w_ce == 0.0:
            return 0
        ig = self.ignore_index if self.ignore_index is not None else -100   # CrossEntropyLoss default
        return _losses.dice_ce(predict, target, ig, w_dice, w_ce, self.reduce_partials)


class symJSDivLoss(torch.nn.Module):
    """loss_functions.py:27-37.  Constructed by the trainers but never called: kept for
    API compatibility as a plain composition of torch ops, it is not part of the hot path."""

    def __init__(self, ):
        super(symJSDivLoss, self).__init__()
        self.KLDivLoss = torch.nn.KLDivLoss()

    def forward(self, predict, target):
        p = predict.softmax(dim=1).clamp(min=1e-10)
        t = target.softmax(dim=1).clamp(min=1e-10)
        return 0.5 * self.KLDivLoss(p.log(), t) + 0.5 * self.KLDivLoss(t.log(), p)


class DiceLoss(torch.nn.Module):
    """loss_functions.py:96-135 (softmax + ignore mask + per-class BinaryDiceLoss(smooth=1, p=2), mean over classes)."""

    def __init__(self, weight=None, num_classes=13, ignore_index=None, **kwargs):
        super(DiceLoss, self).__init__()
        if kwargs:
            raise NotImplementedError("BinaryDiceLoss kwargs other than the defaults (smooth=1, p=2) are not used "
                                      "by any OpenESS trainer")
        if weight is not None:
            raise NotImplementedError("per-class dice weights are not used by any OpenESS trainer "
                                      "(the reference path reads an undefined attribute, loss_functions.py:131)")
        self.kwargs = kwargs
        self.weight = weight
        self.num_classes = num_classes
        self.ignore_index = ignore_index

    def forward(self, predict, target):
        assert predict.shape[1] == self.num_classes, 'predict & target shape do not match'
        ig = self.ignore_index if self.ignore_index is not None else -(1 << 62)
        return _losses.dice_ce(predict, target, ig, 1.0, 0.0)


class NCELoss(torch.nn.Module):
    """loss_functions.py:138-153: PointInfoNCE, CE((k @ q.T) / temperature, arange(M))."""

    def __init__(self, temperature):
        super(NCELoss, self).__init__()
        self.temperature = temperature

    def forward(self, k, q):
        return _losses.infonce(k, q, self.temperature)
