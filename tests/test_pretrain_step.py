"""OpenESS stage-1 pretraining step (row a19): the fused step of openess_b200/training/pretrain_step.py against the
LITERAL formulation of training/pretrain_trainer.py:427-472 + :550-562 + utils/loss_functions.py (restated here with
plain torch ops: sparse one-hot pooling on permuted copies, materialised x_ch256, softmax / one-hot Dice, CE) evaluated on
the same modules.  Every module is separately pinned to a reference golden; this pins the composition and the gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from seeded_weights import seeded_state_dict

pytestmark = pytest.mark.gpu


def _models(dev):
    from types import SimpleNamespace
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.e2vid.model.model import E2VIDRecurrent
    from openess_b200.models.image_model import DilationFeatureExtractor
    from openess_b200.models.style_networks import SemSegE2VID
    z = load_golden("e2vid_tiny")
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = (v == "True") if str(v) in ("True", "False") else (int(v) if str(v).isdigit() else str(v))
    e2vid = E2VIDRecurrent(cfg, latent_only=True)
    e2vid.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}, strict=True)
    e2vid = e2vid.eval().to(dev).fold_bn()
    zs = load_golden("semseg_tiny")
    back = SemSegE2VID(input_c=32, output_c=int(zs["K"]), skip_connect=True, skip_type='concat', text_embeddings_path=None)
    back.load_state_dict({k[4:]: torch.from_numpy(zs[k]) for k in zs.files if k.startswith("sd__")}, strict=True)
    teacher = DilationFeatureExtractor()
    teacher.load_state_dict(seeded_state_dict(teacher, 77), strict=True)
    opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
    return e2vid, back.to(dev), teacher.to(dev), opts, int(zs["K"])


def _literal_step(e2vid, back, teacher, event, frame, pl, sp, S, steps):
    """pretrain_trainer.py:427-472 with the reference's own torch formulation of every block."""
    feat_frame = teacher(frame)                                                     # :434
    states = None
    for i in range(steps):                                                          # :437-441
        ev = event[:, 5 * i:5 * i + 5]
        nz = ev != 0                                                                # inference_utils.py:77-85
        n = nz.sum()
        mean = ev.sum() / n
        std = torch.sqrt((ev ** 2).sum() / n - mean ** 2)
        ev = nz.float() * (ev - mean) / std
        with torch.no_grad():
            _, states, latent = e2vid(ev, states)
    pred, feat_voxel = back({k: v.detach() for k, v in latent.items()})            # :551-553
    logits = pred[1]
    ce = F.cross_entropy(logits, pl, ignore_index=255)                              # loss_functions.py:17-24
    mask = (pl != 255)
    onehot = F.one_hot((pl * mask).long(), logits.shape[1]).permute(0, 3, 1, 2).float() * mask[:, None]
    prob = logits.softmax(1) * mask[:, None]
    dice = 0
    for c in range(logits.shape[1]):                                                # loss_functions.py:80-90, 114-135
        num = 2 * (prob[:, c] * onehot[:, c]).sum() + 1
        den = (prob[:, c] ** 2 + onehot[:, c] ** 2).sum() + 1
        dice = dice + (1 - num / den)
    loss_dense = dice / logits.shape[1] + ce
    B = feat_voxel.shape[0]
    spx = torch.arange(0, B * S, S, device=sp.device)[:, None, None] + sp          # :446-449
    sI = spx.flatten()
    idx = torch.arange(sI.shape[0], device=sp.device)
    with torch.no_grad():
        one_hot = torch.sparse_coo_tensor(torch.stack((sI, idx), 0), torch.ones(sI.shape[0], device=sp.device))
    cnt = torch.sparse.sum(one_hot, 1).to_dense()[:, None] + 1e-6
    k = (one_hot @ feat_voxel.permute(0, 2, 3, 1).flatten(0, 2)) / cnt             # :456-459
    q = (one_hot @ feat_frame.permute(0, 2, 3, 1).flatten(0, 2)) / cnt             # :461-463
    nce = F.cross_entropy((k @ q.t()) / 0.07, torch.arange(k.shape[0], device=k.device))    # loss_functions.py:147-153
    return nce + loss_dense, nce, loss_dense


def test_pretrain_step_matches_literal_reference_formulation():
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.models import image_model as im
    from openess_b200.training.pretrain_step import OpenESSPretrainStep
    from openess_b200.utils.loss_functions import NCELoss, TaskLoss
    dev = torch.device("cuda:0")
    e2vid, back, teacher, opts, K = _models(dev)
    Bn, H, W, S, steps = 2, 32, 48, 10, 3
    g = torch.Generator().manual_seed(11)
    event = torch.randn(Bn, 5 * steps, H, W, generator=g)
    event[torch.rand(event.shape, generator=g) < 0.6] = 0
    frame = torch.rand(Bn, 3, H, W, generator=g)
    pl = torch.randint(0, K, (Bn, H, W), generator=g)
    pl[torch.rand(pl.shape, generator=g) < 0.03] = 255
    sp = torch.randint(0, S, (Bn, H, W), generator=g)
    event, frame, pl, sp = (t.to(dev) for t in (event, frame, pl, sp))
    sd_back = {k: v.clone() for k, v in back.state_dict().items()}
    sd_teacher = {k: v.clone() for k, v in teacher.state_dict().items()}

    from openess_b200.models import style_networks as sn
    im.USE_TENSOR_CORES = False                       # identical fp32 modules on both sides: pins the composition
    sn.TRAIN_ON_TENSOR_CORES = False
    try:
        teacher.train(); back.train()
        ref_total, ref_nce, ref_dense = _literal_step(e2vid, back, teacher, event, frame, pl, sp, S, steps)
        ref_total.backward()
        ref_grads = {n: p.grad.clone() for n, p in list(back.named_parameters()) + list(teacher.named_parameters())
                     if p.grad is not None}
        for p in list(back.parameters()) + list(teacher.parameters()):
            p.grad = None
        back.load_state_dict(sd_back); teacher.load_state_dict(sd_teacher)        # undo BN running-stat updates

        rec = ImageReconstructor(e2vid, H, W, 5, dev, opts)
        step = OpenESSPretrainStep(rec, back, teacher, TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255),
                                   NCELoss(temperature=0.07), nr_events_data_b=steps, superpixel_size=S, lr_voxel=1e-3, lr_frame=1e-3)
        total, losses, _ = step.task_train_step((event, None, frame, pl, sp))
        assert float(losses["contrastive_nce_loss"]) == pytest.approx(float(ref_nce), rel=2e-4)
        assert float(losses["dense_clip_loss"]) == pytest.approx(float(ref_dense), rel=2e-4)
        assert float(total) == pytest.approx(float(ref_total), rel=2e-4)
        total.backward()
        checked = 0
        for n, p in list(back.named_parameters()) + list(teacher.named_parameters()):
            if n in ref_grads:
                assert p.grad is not None, n
                ref = ref_grads[n]
                atol = 2e-3 * float(ref.abs().max()) + 1e-7
                if n.endswith(".model.0.bias") or n.endswith(".model.3.bias"):
                    continue                          # bias in front of an affine-free InstanceNorm: gradient is round-off noise
                np.testing.assert_allclose(p.grad.cpu().numpy(), ref.cpu().numpy(), atol=atol, err_msg=n)
                checked += 1
            else:
                assert p.grad is None, n              # decoder_scale_5, frozen encoder
        assert checked >= 20
        # full train_step: both AdamW optimisers move exactly the trainable parameters
        for p in list(back.parameters()) + list(teacher.parameters()):
            p.grad = None
        before = {n: p.detach().clone() for n, p in teacher.named_parameters()}
        w_before = back.decoder_ch256[0].weight.detach().clone()
        step.train_step((event, None, frame, pl, sp))
        assert not torch.equal(back.decoder_ch256[0].weight, w_before)
        for n, p in teacher.named_parameters():
            assert torch.equal(p, before[n]) == (not n.startswith("decoder")), n
    finally:
        im.USE_TENSOR_CORES = True
        sn.TRAIN_ON_TENSOR_CORES = True

    # tensor-core teacher and task decoder (the production configuration): same step, losses in the TF32 noise class
    back.load_state_dict(sd_back); teacher.load_state_dict(sd_teacher)
    total_tc, losses_tc, _ = step.task_train_step((event, None, frame, pl, sp))
    assert float(losses_tc["dense_clip_loss"]) == pytest.approx(float(ref_dense), rel=2e-2)
    assert float(losses_tc["contrastive_nce_loss"]) == pytest.approx(float(ref_nce), rel=5e-2)


def test_pretrain_step_raw_event_slab_equals_dense_event_tensor():
    """GPU-side sample assembly: raw DSEC records -> the same dense [B, 20*5, H, W] tensor the reference's loader builds."""
    from openess_b200 import voxel
    from openess_b200.training.pretrain_step import OpenESSPretrainStep, RawEvents
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    Hs, Ws, crop, C, Bn, steps, n = 40, 64, 32, 5, 2, 3, 500
    F_ = Bn * steps
    x = torch.from_numpy(rng.integers(0, Ws, F_ * n).astype(np.uint16))
    y = torch.from_numpy(rng.integers(0, Hs, F_ * n).astype(np.uint16))
    t = torch.from_numpy(np.concatenate([np.sort(rng.integers(0, 50000, n)) + 1000 + 50000 * f for f in range(F_)]).astype(np.int64))
    p = torch.from_numpy(rng.integers(0, 2, F_ * n).astype(np.uint8))
    yy, xx = np.meshgrid(np.arange(Hs), np.arange(Ws), indexing="ij")
    rmap = torch.from_numpy((np.stack([xx, yy], -1) + rng.uniform(-0.7, 0.7, (Hs, Ws, 2))).astype(np.float32))
    fo = torch.arange(0, (F_ + 1) * n, n, dtype=torch.int64)
    step = OpenESSPretrainStep.__new__(OpenESSPretrainStep)
    step.device, step.input_channels_b, step.nr_events_data_b = dev, C, steps
    dense = step.event_tensor(RawEvents(x, y, t, p, fo, rmap, (Hs, Ws), crop))
    assert tuple(dense.shape) == (Bn, steps * C, crop, Ws)
    for f in range(F_):                                   # frame by frame through the single-frame mirror of VoxelGrid
        sl = slice(f * n, (f + 1) * n)
        g = voxel.dsec_events_to_voxel_grid(x[sl].to(dev), y[sl].to(dev), t[sl].to(dev), p[sl].to(dev), rmap.to(dev), C)
        b, i = divmod(f, steps)
        assert torch.equal(dense[b, i * C:(i + 1) * C], g[0, :, :crop])


def test_pretrain_step_matches_reference_trainer_golden():
    """a19 against the REFERENCE's own trainer: tests/golden/pretrain_step.npz holds losses, every gradient and the parameters
    after one optimiser step of `OpenESSPretrainModel.task_train_step` / `train_step` (training/pretrain_trainer.py:324-361,
    364-372, 427-472, 550-562), executed unmodified on CPU by oracle/make_golden_trainer.py with the reference's module
    classes and the same seeded weights."""
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.models import image_model as im
    from openess_b200.models import style_networks as sn
    from openess_b200.training.pretrain_step import OpenESSPretrainStep
    from openess_b200.utils.loss_functions import NCELoss, TaskLoss
    z = load_golden("pretrain_step")
    dev = torch.device("cuda:0")
    e2vid, back, teacher, opts, K = _models(dev)
    assert K == int(z["K"])
    S, steps, stride = int(z["S"]), int(z["steps"]), int(z["stride"])
    event, frame, pl, sp = (torch.from_numpy(z[k]).to(dev) for k in ("event", "frame", "pl", "sp"))
    H, W = event.shape[-2:]
    sd_back = {k: v.clone() for k, v in back.state_dict().items()}
    sd_teacher = {k: v.clone() for k, v in teacher.state_dict().items()}
    im.USE_TENSOR_CORES = False                       # fp32 modules: the golden is the reference's fp32 CPU run
    sn.TRAIN_ON_TENSOR_CORES = False
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False           # torch's default TF32 convolutions drift by percents through the 16 conv +
    try:                                              # InstanceNorm layers of the (x8-weight-scaled) tiny SemSegE2VID trunk
        rec = ImageReconstructor(e2vid, H, W, 5, dev, opts)
        step = OpenESSPretrainStep(rec, back, teacher, TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255),
                                   NCELoss(temperature=0.07), nr_events_data_b=steps, superpixel_size=S, lr_voxel=1e-3, lr_frame=1e-3)
        total, losses, outputs = step.task_train_step((event, None, frame, pl, sp))
        assert float(losses["contrastive_nce_loss"]) == pytest.approx(float(z["nce"]), rel=3e-4)
        assert float(losses["dense_clip_loss"]) == pytest.approx(float(z["dense"]), rel=3e-4)
        assert float(total) == pytest.approx(float(z["total"]), rel=3e-4)
        if z["logits"].size:                          # the reference's frame2voxel branch leaves `outputs` empty
            np.testing.assert_allclose(outputs["pred"][1].detach().cpu().numpy(), z["logits"], atol=2e-3 * float(np.abs(z["logits"]).max()))
        total.backward()
        named = {"back_end." + n: p for n, p in back.named_parameters()}
        named.update({"model_frame." + n: p for n, p in teacher.named_parameters()})
        nograd = set(str(n) for n in z["nograd"])
        checked = 0
        for n, p in named.items():
            if n in nograd:
                assert p.grad is None, n              # decoder_scale_5, frozen ResNet-50 encoder
                continue
            ref = z["grad__" + n]
            got = p.grad.cpu().numpy()
            if got.size != ref.size:
                got = got.reshape(-1)[::stride]
            if n.endswith(".model.0.bias") or n.endswith(".model.3.bias"):
                continue                              # bias in front of an affine-free InstanceNorm: gradient is round-off noise
            got = got.reshape(ref.shape)
            if n.startswith(("back_end.decoder_ch", "back_end.text_emb", "model_frame.")):
                np.testing.assert_allclose(got, ref, atol=3e-3 * float(np.abs(ref).max()) + 1e-7, err_msg=n)
            else:
                # trunk of the tiny SemSegE2VID golden (weights scaled x8, up to 16 conv + InstanceNorm layers): CPU and GPU
                # fp32 summation orders drift by ~1 % of the largest entry on the way back (the forward losses agree to 3e-4)
                err = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12)
                assert err < 5e-2, (n, err)
            checked += 1
        assert checked >= 20
        # one full train_step from the same starting point: zero_grad + forward + backward + 2 x AdamW
        back.load_state_dict(sd_back); teacher.load_state_dict(sd_teacher)
        for p in named.values():
            p.grad = None
        step = OpenESSPretrainStep(rec, back, teacher, step.task_loss, step.nce_loss, nr_events_data_b=steps, superpixel_size=S,
                                   lr_voxel=1e-3, lr_frame=1e-3)
        _, _, final = step.train_step((event, None, frame, pl, sp))
        assert float(final) == pytest.approx(float(z["step_total"]), rel=3e-4)
        after = [k for k in z.files if k.startswith("after__")]
        assert len(after) >= 4
        for key in after:
            n = key[len("after__"):]
            ref = z[key]
            got = named[n].detach().cpu().numpy()
            if got.size != ref.size:
                got = got.reshape(-1)[::stride]
            d = np.abs(got.reshape(ref.shape) - ref)
            # AdamW's first step moves every weight by ~lr * sign(g): entries whose gradient is round-off noise may flip sign
            assert float(d.max()) <= 2.1e-3, n
            assert float((d < 2e-5).mean()) > 0.98, (n, float((d < 2e-5).mean()))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32_was
        im.USE_TENSOR_CORES = True
        sn.TRAIN_ON_TENSOR_CORES = True


@pytest.mark.gpu
def test_device_prefetcher_moves_nested_batches_on_a_side_stream():
    """training/prefetch.py: the batch fed last comes back on the device with equal contents (RawEvents slab, dense tensors, None
    and nested tuples kept), ordered after the copies; take() without a feed raises."""
    import numpy as np
    from openess_b200.training.prefetch import DevicePrefetcher
    from openess_b200.training.pretrain_step import RawEvents
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    n = 50_000
    ev = RawEvents(torch.randint(0, 640, (n,), generator=g).to(torch.uint16).pin_memory(),
                   torch.randint(0, 480, (n,), generator=g).to(torch.uint16).pin_memory(),
                   torch.arange(n, dtype=torch.int64).to(torch.uint32).pin_memory(),
                   torch.randint(0, 2, (n,), generator=g).to(torch.uint8).pin_memory(),
                   torch.tensor([0, n // 2, n], dtype=torch.int64), torch.rand(480, 640, 2, generator=g), (480, 640), 440)
    frame = torch.rand(2, 3, 8, 8, generator=g).pin_memory()
    batch = (ev, None, frame, (torch.arange(6).pin_memory(), "keep"))
    pf = DevicePrefetcher(dev)
    with pytest.raises(RuntimeError):
        pf.take()
    for _ in range(3):                                       # re-feeding recycles the buffers of the batch taken before
        pf.feed(batch)
        cur = pf.take()
        assert isinstance(cur[0], RawEvents) and cur[1] is None and cur[3][1] == "keep"
        assert cur[0].x.is_cuda and cur[2].is_cuda and cur[3][0].is_cuda and cur[0].sensor_hw == (480, 640) and cur[0].crop_h == 440
        torch.cuda.synchronize()
        for a, b in ((cur[0].x, ev.x), (cur[0].y, ev.y), (cur[0].t, ev.t), (cur[0].p, ev.p), (cur[0].frame_offsets, ev.frame_offsets),
                     (cur[0].rectify_map, ev.rectify_map), (cur[2], frame), (cur[3][0], batch[3][0])):
            assert np.array_equal(a.cpu().numpy(), b.numpy())


@pytest.mark.gpu
def test_encoder_loop_cuda_graph_equals_eager():
    """training/graphs.py: the frozen recurrent encoder loop captured as one CUDA graph (third call on) gives the same latents as
    the eager loop, for changing inputs in the persistent event buffer, and is re-captured after reset()."""
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    from openess_b200.training.graphs import GraphedEncoderLoop
    dev = torch.device("cuda:0")
    e2vid, _, _, opts, _ = _models(dev)
    Bn, H, W, steps, C = 2, 32, 48, 3, 5
    rec = ImageReconstructor(e2vid, H, W, C, dev, opts)
    loop = GraphedEncoderLoop(rec, steps, C)
    ref = GraphedEncoderLoop(ImageReconstructor(e2vid, H, W, C, dev, opts), steps, C)
    ref.disabled = True                                      # always eager
    g = torch.Generator().manual_seed(21)
    buf = torch.empty(Bn, C * steps, H, W, device=dev)       # persistent input buffer (what the voxeliser writes into)
    for it in range(8):
        ev = torch.randn(Bn, C * steps, H, W, generator=g)
        ev[torch.rand(ev.shape, generator=g) < 0.6] = 0
        buf.copy_(ev.to(dev))
        got = loop(buf)
        want = ref(buf.clone())
        assert (loop.graph is not None) == (it in (2, 3, 6, 7))      # two eager calls, capture on the third; again after reset()
        for k in want:
            assert torch.equal(got[k], want[k]), (it, k)
        if it == 3:
            loop.reset()
            assert loop.graph is None
