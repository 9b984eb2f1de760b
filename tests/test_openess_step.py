"""`OpenESSModel` step, branch frame2recon (training/openess_trainer.py:339-358, 360-372, 476-529) against the golden produced by
the REFERENCE's own trainer class on CPU (oracle/make_golden_trainer.py: two reference deeplabv3_resnet50 networks with seeded
weights, losses + selected gradients + parameters after one AdamW step)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from seeded_weights import seeded_state_dict

pytestmark = pytest.mark.gpu


def _net(seed, K, dev):
    from openess_b200.models.deeplabv3 import deeplabv3_resnet50
    m = deeplabv3_resnet50(num_classes=K, text_embeddings_path=None, output_stride=32, pretrained_backbone='')
    m.load_state_dict(seeded_state_dict(m, seed), strict=True)
    m.classifier.ASPP.project[3].p = 0.0                  # Dropout is random: disabled in the golden and here
    return m.to(dev)


def test_openess_frame2recon_step_matches_reference_trainer_golden():
    from openess_b200.training.openess_step import OpenESSStep
    from openess_b200.utils.loss_functions import NCELoss, TaskLoss
    z = load_golden("openess_step")
    dev = torch.device("cuda:0")
    K, stride = int(z["K"]), int(z["stride"])
    model_frame, model_recon = _net(4, K, dev), _net(5, K, dev)
    batch = tuple(None if k is None else torch.from_numpy(z[k]) for k in ("frame", None, "recon", "pl", "sp"))
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False               # the golden is the reference's fp32 CPU run
    try:
        sd = [{k: v.clone() for k, v in m.state_dict().items()} for m in (model_frame, model_recon)]
        step = OpenESSStep(model_frame, model_recon, TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255),
                           NCELoss(temperature=0.07), lr_recon=1e-3, lr_frame=1e-3)
        total, losses, _ = step.task_train_step(batch)
        for k in ("semseg_frame_loss", "semseg_recon_loss", "cons_feat_loss", "cons_pred_loss", "contrastive_nce_loss"):
            # the InfoNCE logits are products of UNNORMALISED 256-channel features divided by T = 0.07 (values of several
            # hundred, loss ~ 32): CPU / GPU fp32 differences of 1e-5 in the features show up as 1e-3 of that loss
            assert float(losses[k]) == pytest.approx(float(z["loss__" + k]), rel=3e-3 if k == "contrastive_nce_loss" else 5e-4), k
        assert float(total) == pytest.approx(float(z["total"]), rel=3e-3)
        total.backward()
        checked = 0
        for prefix, m in (("model_frame.", model_frame), ("model_recon.", model_recon)):
            named = dict(m.named_parameters())
            assert sorted(n for n, p in named.items() if p.grad is None) == sorted(str(n) for n in z["nograd__" + prefix])
            for key in z.files:
                if not key.startswith("grad__" + prefix):
                    continue
                n = key[len("grad__" + prefix):]
                ref = z[key]
                got = named[n].grad.cpu().numpy()
                if got.size != ref.size:
                    got = got.reshape(-1)[::stride]
                # train-mode BatchNorm over a batch of 2 x 4 x 6 positions amplifies fp32 summation-order differences on the way back
                # through 50 layers (measured: 3e-2 at backbone.conv1, < 1e-2 in the head)
                err = np.linalg.norm(got.reshape(ref.shape) - ref) / (np.linalg.norm(ref) + 1e-12)
                assert err < 6e-2, (prefix + n, err)
                checked += 1
        assert checked == 10
        for m, s0 in zip((model_frame, model_recon), sd):
            m.load_state_dict(s0)
            for p in m.parameters():
                p.grad = None
        step = OpenESSStep(model_frame, model_recon, step.task_loss, step.nce_loss, lr_recon=1e-3, lr_frame=1e-3)
        _, _, final = step.train_step(batch)
        assert float(final) == pytest.approx(float(z["step_total"]), rel=3e-3)
        for prefix, m in (("model_frame.", model_frame), ("model_recon.", model_recon)):
            named = dict(m.named_parameters())
            for key in z.files:
                if key.startswith("after__" + prefix):
                    ref = z[key]
                    got = named[key[len("after__" + prefix):]].detach().cpu().numpy()
                    if got.size != ref.size:
                        got = got.reshape(-1)[::stride]
                    d = np.abs(got.reshape(ref.shape) - ref)
                    assert float(d.max()) <= 2.1e-3 and float((d < 5e-5).mean()) > 0.97, key
    finally:
        torch.backends.cudnn.allow_tf32 = tf32_was
