"""Drop-in mirror of the reference's evaluation/metrics.py; the confusion matrix is built by a CUDA kernel
(integer exact) and accumulated on the device (the reference moves it to the CPU every batch)."""
import torch

from .. import losses as _losses


def semseg_compute_confusion(y_hat_lbl, y_lbl, num_classes, ignore_label):
    """metrics.py:4-23 -> int64 [K, K] with conf[gt, pred]."""
    assert torch.is_tensor(y_hat_lbl) and torch.is_tensor(y_lbl), 'Inputs must be torch tensors'
    assert y_lbl.device == y_hat_lbl.device, 'Input tensors have different device placement'
    assert y_hat_lbl.dim() == 3 or y_hat_lbl.dim() == 4 and y_hat_lbl.shape[1] == 1
    assert y_lbl.dim() == 3 or y_lbl.dim() == 4 and y_lbl.shape[1] == 1
    src = y_lbl.device
    if not torch.cuda.is_available():
        raise RuntimeError("openess_b200 metrics need a CUDA device (no CPU fallback)")
    dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    conf = _losses.confusion(y_hat_lbl.to(dev), y_lbl.to(dev), num_classes, ignore_label, status=status)
    assert int(status.item()) == 0, 'Internal error'      # metrics.py:21 (bincount larger than K*K)
    return conf.to(src)


def semseg_accum_confusion_to_iou(confusion_accum):
    """metrics.py:26-30 (float64)."""
    conf = confusion_accum.double()
    diag = conf.diag()
    iou_per_class = 100 * diag / (conf.sum(dim=1) + conf.sum(dim=0) - diag).clamp(min=1e-12)
    iou_mean = iou_per_class.mean()
    return iou_mean, iou_per_class


def semseg_accum_confusion_to_acc(confusion_accum):
    """metrics.py:33-36 (float64)."""
    conf = confusion_accum.double()
    diag = conf.diag()
    acc = 100 * diag.sum() / (conf.sum(dim=1).sum()).clamp(min=1e-12)
    return acc


class MetricsSemseg:
    """metrics.py:39-67."""

    def __init__(self, num_classes, ignore_label, class_names):
        self.num_classes = num_classes
        self.ignore_label = ignore_label
        self.class_names = class_names
        self.metrics_acc = None

    def reset(self):
        self.metrics_acc = None

    def update_batch(self, y_hat_lbl, y_lbl):
        with torch.no_grad():
            metrics_batch = semseg_compute_confusion(y_hat_lbl, y_lbl, self.num_classes, self.ignore_label).cpu()
            if self.metrics_acc is None:
                self.metrics_acc = metrics_batch
            else:
                self.metrics_acc += metrics_batch.to(self.metrics_acc.device)     # the accumulator may live on the device

    def update_batch_logits(self, logits, y_lbl):
        """Fused validation step: `update_batch(logits.argmax(dim=1), y_lbl)` of base_trainer_ov.py:463-471
        in one kernel, accumulated on the DEVICE (no per-batch .cpu() round trip); get_metrics_summary() is unchanged."""
        with torch.no_grad():
            if self.metrics_acc is None or not self.metrics_acc.is_cuda:
                prev = self.metrics_acc
                self.metrics_acc = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=logits.device)
                if prev is not None:
                    self.metrics_acc += prev.to(logits.device)
            _losses.argmax_confusion(logits, y_lbl.to(logits.device), self.ignore_label, out=self.metrics_acc)

    def get_metrics_summary(self):
        if self.metrics_acc is not None and self.metrics_acc.is_cuda:
            self.metrics_acc = self.metrics_acc.cpu()       # the float64 summary is computed on the host, like the reference
        iou_mean, iou_per_class = semseg_accum_confusion_to_iou(self.metrics_acc)
        out = {self.class_names[i]: iou for i, iou in enumerate(iou_per_class)}
        out['miou'] = iou_mean
        acc = semseg_accum_confusion_to_acc((self.metrics_acc))
        out['acc'] = acc
        out['cm'] = self.metrics_acc
        return out
