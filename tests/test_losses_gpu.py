"""GPU parity tests of the loss-side kernels (segment pool, InfoNCE, Dice+CE, confusion, normalise) through the
reference-mirroring modules.  Golden vectors come from the reference's own torch code; fresh inputs are checked
against the float64 oracle.  Tolerances: float32 kernels vs float64 oracle / torch float32: rtol 2e-5 on
losses, 2e-4 on gradients (the reference's float32 reductions carry the same order of error); confusion
matrices and mIoU are integer-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_pool_nce_golden_forward_backward(dev):
    from openess_b200.training.superpixel import pooled_pair
    from openess_b200.utils.loss_functions import NCELoss
    z = load_golden("losses")
    fv = torch.from_numpy(z["pool__feat_voxel"]).to(dev).requires_grad_(True)
    ff = torch.from_numpy(z["pool__feat_frame"]).to(dev).requires_grad_(True)
    sp = torch.from_numpy(z["pool__superpixels"]).to(dev)
    k, q = pooled_pair(fv, ff, sp, int(z["pool__S"]))
    assert tuple(k.shape) == z["pool__k"].shape          # M = max id' + 1, incl. the aliasing ids >= S
    np.testing.assert_allclose(k.detach().cpu().numpy(), z["pool__k"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(q.detach().cpu().numpy(), z["pool__q"], rtol=2e-5, atol=2e-6)
    loss = NCELoss(0.07)(k, q)
    assert float(loss) == pytest.approx(float(z["pool__nce"]), rel=2e-5)
    loss.backward()
    np.testing.assert_allclose(fv.grad.cpu().numpy(), z["pool__d_feat_voxel"], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(ff.grad.cpu().numpy(), z["pool__d_feat_frame"], rtol=2e-4, atol=2e-7)


def test_task_loss_golden_forward_backward(dev):
    from openess_b200.utils.loss_functions import DiceLoss, TaskLoss
    z = load_golden("losses")
    logits = torch.from_numpy(z["task__logits"]).to(dev).requires_grad_(True)
    target = torch.from_numpy(z["task__target"]).to(dev)
    K = logits.shape[1]
    tl = TaskLoss(losses=["dice", "cross_entropy"], num_classes=K, ignore_index=255)
    loss = tl(logits, target)
    assert float(loss) == pytest.approx(float(z["task__total"]), rel=2e-5)
    (loss * 1.5).backward()
    np.testing.assert_allclose(logits.grad.cpu().numpy(), 1.5 * z["task__dlogits"], rtol=2e-4, atol=2e-8)
    dice = DiceLoss(num_classes=K, ignore_index=255)(logits.detach(), target)
    assert float(dice) == pytest.approx(float(z["task__dice"]), rel=2e-5)
    ce_only = TaskLoss(losses=["cross_entropy"], num_classes=K, ignore_index=255)(logits.detach(), target)
    assert float(ce_only) == pytest.approx(float(z["task__total"]) - float(z["task__dice"]), rel=5e-5)


def test_metrics_golden_integer_exact(dev):
    from openess_b200.evaluation import metrics
    z = load_golden("losses")
    pred, gt = torch.from_numpy(z["met__pred"]), torch.from_numpy(z["met__gt"])
    conf = metrics.semseg_compute_confusion(pred.to(dev), gt.to(dev), 11, 255)
    assert conf.device == dev and conf.dtype == torch.int64
    assert np.array_equal(conf.cpu().numpy(), z["met__conf"])
    ms = metrics.MetricsSemseg(11, 255, [str(i) for i in range(11)])
    ms.update_batch(pred[:2], gt[:2])                       # CPU tensors in, like valBatchStep
    ms.update_batch(pred[2:].unsqueeze(1), gt[2:].unsqueeze(1))   # [B,1,H,W] accepted (metrics.py:8-13)
    s = ms.get_metrics_summary()
    assert float(s["miou"]) == pytest.approx(float(z["met__miou"]), rel=1e-12)
    assert float(s["acc"]) == pytest.approx(float(z["met__acc"]), rel=1e-12)
    assert np.array_equal(s["cm"].numpy(), z["met__conf"])
    with pytest.raises(AssertionError):                     # bincount beyond K*K -> the reference's assert
        metrics.semseg_compute_confusion(pred.to(dev) + 200, gt.to(dev), 11, 255)


@pytest.mark.parametrize("B,Cf,H,W,S", [(2, 32, 24, 36, 12), (3, 256, 40, 64, 25), (1, 7, 15, 17, 5), (4, 64, 110, 160, 100)])
def test_segpool_vs_oracle(dev, oracle, B, Cf, H, W, S):
    from openess_b200 import losses
    rng = np.random.default_rng(B * 100 + Cf)
    feat = rng.normal(0, 1, (B, Cf, H, W)).astype(np.float32)
    # blocky superpixels (runs along rows, like SAM / SLIC maps) + a few ids >= S (aliasing quirk)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    seg = ((yy // max(H // 5, 1)) * 5 + xx // max(W // 5, 1)) % S
    seg = np.stack([np.roll(seg, b, 1) for b in range(B)]).astype(np.int64)
    seg[0, 0, :3] = S + 1
    M = int((seg + np.arange(B)[:, None, None] * S).max()) + 1
    ref_p, ref_c = oracle.segpool(feat, seg, S, M)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    p, c = losses.segpool_forward(torch.from_numpy(feat).to(dev), torch.from_numpy(seg).to(dev), S, M, status=st)
    assert int(st.item()) == 0
    np.testing.assert_array_equal(c.cpu().numpy(), ref_c)
    np.testing.assert_allclose(p.cpu().numpy(), ref_p, rtol=2e-5, atol=2e-6)
    # backward: d_feat[b, c, pix] = d_pooled[id', c] / (count + 1e-6)
    dp = rng.normal(0, 1, (M, Cf)).astype(np.float32)
    d = losses.segpool_backward(torch.from_numpy(dp).to(dev), torch.from_numpy(seg).to(dev), c, S, (B, Cf, H, W))
    ids = seg + np.arange(B)[:, None, None] * S
    want = (dp / (ref_c[:, None] + np.float32(1e-6)))[ids]          # [B,H,W,Cf]
    np.testing.assert_allclose(d.cpu().numpy(), want.transpose(0, 3, 1, 2), rtol=1e-6, atol=1e-7)
    # out-of-range ids are skipped and flagged
    bad = seg.copy()
    bad[0, 1, 1] = M + 5
    losses.segpool_forward(torch.from_numpy(feat).to(dev), torch.from_numpy(bad).to(dev), S, M, status=st)
    assert int(st.item()) == 1


@pytest.mark.parametrize("M,D", [(50, 32), (200, 256), (801, 256), (130, 64), (3200, 256), (1027, 256), (2050, 64)])
def test_infonce_vs_oracle(dev, oracle, M, D):
    from openess_b200 import losses
    rng = np.random.default_rng(M)
    k = rng.normal(0, 1, (M, D)).astype(np.float32)
    q = (0.7 * k + 0.5 * rng.normal(0, 1, (M, D))).astype(np.float32)
    k /= np.linalg.norm(k, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    k[3] = 0                                                         # empty superpixel rows stay (Appendix B.9)
    big = M > 1000
    if big:
        kt, qt = torch.from_numpy(k).double(), torch.from_numpy(q).double()
        kt.requires_grad_(True); qt.requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(kt @ qt.T / 0.07, torch.arange(M))
        ref.backward()
        ref_loss, rdk, rdq = float(ref), kt.grad.numpy(), qt.grad.numpy()
    else:
        ref_loss, rdk, rdq = oracle.infonce(k, q, 0.07, grad=True)
    kd = torch.from_numpy(k).to(dev).requires_grad_(True)
    qd = torch.from_numpy(q).to(dev).requires_grad_(True)
    loss = losses.infonce(kd, qd, 0.07)
    # M >= 1024: tensor-core path ("3xTF32" operand splitting, fp32 accumulate); the loss is the mean of lse_i - s_ii with both
    # terms of magnitude 1 / T = 14.3, i.e. one fp32 ulp there (9.5e-7) is 2e-5 of a loss of 0.05
    assert float(loss) == pytest.approx(ref_loss, rel=4e-5 if M >= 1024 else 2e-5)
    loss.backward()
    np.testing.assert_allclose(kd.grad.cpu().numpy(), rdk, rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(qd.grad.cpu().numpy(), rdq, rtol=2e-4, atol=2e-7)


@pytest.mark.parametrize("B,K,H,W", [(2, 11, 33, 47), (1, 6, 200, 352), (3, 19, 20, 20), (2, 40, 16, 16)])
def test_dice_ce_vs_oracle(dev, oracle, B, K, H, W):
    from openess_b200 import losses
    rng = np.random.default_rng(K)
    logits = rng.normal(0, 3, (B, K, H, W)).astype(np.float32)
    target = rng.integers(0, K, (B, H, W)).astype(np.int64)
    target[rng.random(target.shape) < 0.03] = 255
    ref = oracle.dice_ce(logits, target, 255, grad=True)
    ld = torch.from_numpy(logits).to(dev).requires_grad_(True)
    loss = losses.dice_ce(ld, torch.from_numpy(target).to(dev), 255)
    assert float(loss) == pytest.approx(ref["total"], rel=2e-5)
    loss.backward()
    np.testing.assert_allclose(ld.grad.cpu().numpy(), ref["dlogits"], rtol=3e-4, atol=2e-9)
    # global-batch semantics: partial sums of two half-batches add up to the full-batch partials
    if B >= 2:
        full = losses.dice_ce_partials(ld.detach(), torch.from_numpy(target).to(dev), 255)
        a = losses.dice_ce_partials(ld.detach()[:1].contiguous(), torch.from_numpy(target[:1]).to(dev), 255)
        b = losses.dice_ce_partials(ld.detach()[1:].contiguous(), torch.from_numpy(target[1:]).to(dev), 255)
        torch.testing.assert_close(a + b, full, rtol=1e-9, atol=1e-9)


def test_confusion_vs_oracle_fullsize(dev, oracle):
    from openess_b200 import losses
    rng = np.random.default_rng(5)
    pred = rng.integers(0, 11, (8, 440, 640)).astype(np.int64)
    gt = rng.integers(0, 11, (8, 440, 640)).astype(np.int64)
    gt[rng.random(gt.shape) < 0.1] = 255
    ref = oracle.confusion(pred, gt, 11, 255)
    conf = losses.confusion(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev), 11, 255)
    assert np.array_equal(conf.cpu().numpy(), ref)
    conf2 = losses.confusion(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev), 11, 255, out=conf)
    assert np.array_equal(conf2.cpu().numpy(), 2 * ref)            # accumulates
    assert int(conf2.sum()) == 2 * int((gt != 255).sum())          # checksum: every valid pixel counted once
    big = losses.confusion(torch.from_numpy(pred % 7).to(dev) * 10, torch.from_numpy(np.where(gt == 255, 255, gt * 9)).to(dev),
                           100, 255)                                # K*K > smem bins -> global-atomic path
    assert int(big.sum()) == int((gt != 255).sum())


def test_event_preprocessor_matches_oracle(dev, oracle):
    from types import SimpleNamespace
    from openess_b200.e2vid.utils.inference_utils import EventPreprocessor
    rng = np.random.default_rng(9)
    x = rng.normal(0, 1.2, (2, 5, 56, 72)).astype(np.float32)
    x[rng.random(x.shape) < 0.75] = 0
    ref, stats = oracle.nonzero_standardize(x)
    pre = EventPreprocessor(SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False))
    xin = torch.from_numpy(x).to(dev)
    out = pre(xin)
    assert out.data_ptr() != xin.data_ptr() and torch.equal(xin.cpu(), torch.from_numpy(x))   # input untouched
    assert np.array_equal(out.cpu().numpy() == 0, ref == 0)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=2e-6)
    # two-phase path with an (identity) cross-rank reduction hook gives the same result
    pre.reduce_stats = lambda s: s
    np.testing.assert_allclose(pre(xin).cpu().numpy(), out.cpu().numpy(), rtol=0, atol=0)
    z = torch.zeros(1, 5, 8, 8, device=dev)
    assert not pre(z).any()                                          # num_nonzeros == 0 -> unchanged


@pytest.mark.gpu
def test_argmax_confusion_fused_vs_oracle_and_metrics(dev, oracle):
    """Fused validation step: argmax + confusion in one pass == oracle.confusion(argmax(logits), gt),
    integer exact, including ties (first maximum wins, like torch.argmax / np.argmax) and ignore pixels."""
    from openess_b200 import losses
    from openess_b200.evaluation.metrics import MetricsSemseg
    rng = np.random.default_rng(9)
    B, K, H, W = 3, 11, 88, 128
    logits = rng.normal(0, 1, (B, K, H, W)).astype(np.float32)
    logits[:, 3] = np.where(rng.random((B, H, W)) < 0.2, logits[:, 7], logits[:, 3])      # exact ties between classes 3 and 7
    gt = rng.integers(0, K, (B, H, W)).astype(np.int64)
    gt[rng.random(gt.shape) < 0.1] = 255
    ref = oracle.confusion(np.argmax(logits, 1).astype(np.int64), gt, K, 255)
    lg, g = torch.from_numpy(logits).to(dev), torch.from_numpy(gt).to(dev)
    conf = losses.argmax_confusion(lg, g, 255)
    assert np.array_equal(conf.cpu().numpy(), ref)
    names = [str(i) for i in range(K)]
    m1, m2 = MetricsSemseg(K, 255, names), MetricsSemseg(K, 255, names)
    for _ in range(2):
        m1.update_batch(lg.argmax(dim=1), g)                   # reference call sequence (base_trainer_ov.py:463-471)
        m2.update_batch_logits(lg, g)                          # fused, accumulated on the device
    s1, s2 = m1.get_metrics_summary(), m2.get_metrics_summary()
    assert np.array_equal(s1['cm'].cpu().numpy(), s2['cm'].cpu().numpy()) and np.array_equal(s2['cm'].cpu().numpy(), 2 * ref)
    assert float(s1['miou']) == float(s2['miou']) and float(s1['acc']) == float(s2['acc'])


def test_task_loss_global_partials_gradient_scaling(dev, monkeypatch):
    """`TaskLoss.reduce_partials` (all-reduced Dice / CE partial sums) combined with the AVERAGED gradient
    all-reduce.  Two ranks are emulated on one device: each shard's partial sums are completed with the other shard's, each
    shard back-propagates, and the averaged input gradients must equal the gradient of the single-process loss on the
    concatenated batch (exact global-batch semantics, SURVEY.md 8e option B)."""
    from openess_b200 import losses, parallel
    from openess_b200.utils.loss_functions import TaskLoss
    g = torch.Generator().manual_seed(5)
    K, H, W = 7, 24, 40
    logits = [torch.randn(2, K, H, W, generator=g).to(dev).requires_grad_(True) for _ in range(2)]
    target = [torch.randint(0, K, (2, H, W), generator=g).to(dev) for _ in range(2)]
    for t in target:
        t[torch.rand(t.shape, device=dev) < 0.05] = 255
    full_logits = torch.cat([l.detach() for l in logits]).requires_grad_(True)
    tl = TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    full = tl(full_logits, torch.cat(target))
    full.backward()
    world = 2
    monkeypatch.setattr(parallel, "world_size", lambda: world)
    parts = [losses.dice_ce_partials(l.detach().contiguous(), t.contiguous(), 255) for l, t in zip(logits, target)]
    for r in range(world):
        other = parts[1 - r]
        tl_r = TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
        tl_r.reduce_partials = lambda p, other=other: p.add_(other)          # what parallel.allreduce_sum_ does across ranks
        loss_r = tl_r(logits[r], target[r])
        assert float(loss_r) == pytest.approx(float(full), rel=1e-5)         # every rank sees the global loss
        loss_r.backward()
    # the data-parallel step AVERAGES gradients over ranks; the parameter gradient is linear in d(loss)/d(logits)
    got = torch.cat([l.grad for l in logits]) / world
    torch.testing.assert_close(got, full_logits.grad, rtol=2e-4, atol=1e-9)


def _reference_task_loss(logits, target, ignore, K):
    """utils/loss_functions.py:96-135 DiceLoss (class `ignore_index` skipped, sum / K) + CrossEntropyLoss(ignore_index), torch."""
    import torch.nn.functional as F
    mask = (target != ignore)
    t = target.clone()
    t[~mask] = 0
    onehot = F.one_hot(t, K).permute(0, 3, 1, 2).double() * mask[:, None].double()
    p = torch.softmax(logits.double(), 1) * mask[:, None].double()
    total = 0.0
    for c in range(K):
        if c != ignore:
            num = 2.0 * (p[:, c] * onehot[:, c]).sum() + 1.0
            den = (p[:, c] ** 2 + onehot[:, c] ** 2).sum() + 1.0
            total = total + (1.0 - num / den)
    ce = F.cross_entropy(logits.double(), target, ignore_index=ignore)
    return total / K + ce


def test_dice_ce_ignore_index_inside_class_range(dev):
    """with ignore_index in [0, K) the reference skips that CLASS's Dice term (still dividing by K) and ignores
    those pixels; loss and gradient against a float64 torch statement of loss_functions.py:96-135."""
    from openess_b200 import losses
    g = torch.Generator().manual_seed(3)
    B, K, H, W, ig = 2, 7, 19, 23, 2
    logits = (torch.randn(B, K, H, W, generator=g) * 2).to(dev).requires_grad_(True)
    target = torch.randint(0, K, (B, H, W), generator=g).to(dev)
    loss = losses.dice_ce(logits, target, ig)
    ref_in = logits.detach().clone().requires_grad_(True)
    ref = _reference_task_loss(ref_in, target, ig, K)
    assert float(loss) == pytest.approx(float(ref), rel=2e-5)
    loss.backward()
    ref.backward()
    np.testing.assert_allclose(logits.grad.cpu().numpy(), ref_in.grad.cpu().numpy(), rtol=3e-4, atol=2e-8)
    losses.label_check.check()


def test_dice_ce_out_of_range_labels_are_reported(dev):
    """a target that is neither a class nor ignore_index makes the reference raise; the fused kernel counts it and the
    host raises at the next call (or on `label_check.check()`), without a synchronisation in the step itself."""
    from openess_b200 import losses
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(1, 5, 8, 8, generator=g).to(dev)
    target = torch.randint(0, 5, (1, 8, 8), generator=g).to(dev)
    losses.label_check.check()
    losses.dice_ce(logits, target, 255)
    losses.label_check.check()                               # clean labels: nothing raised
    target[0, 3, 3] = 77
    losses.dice_ce(logits, target, 255)
    with pytest.raises(ValueError, match="outside"):
        losses.label_check.check()
    losses.label_check.check()                               # reported once


def test_batchnorm_momentum_none_is_cumulative_average(dev):
    """torch.nn.BatchNorm2d(momentum=None) uses the cumulative moving average 1 / num_batches_tracked."""
    from openess_b200 import ops
    g = torch.Generator().manual_seed(6)
    bn = torch.nn.BatchNorm2d(16, momentum=None).to(dev).train()
    ref = torch.nn.BatchNorm2d(16, momentum=None).to(dev).train()
    for i in range(3):
        x = (torch.randn(2, 16, 9, 11, generator=g) * (i + 1) + i).to(dev).contiguous(memory_format=torch.channels_last)
        want = ref(x)
        got = ops.batchnorm_nhwc_(x.clone(memory_format=torch.channels_last), bn)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-4, atol=1e-6)
    assert int(bn.num_batches_tracked) == 3
