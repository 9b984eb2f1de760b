"""Golden vectors for rows a19 and a9 produced by the REFERENCE's own trainer / reconstructor classes on CPU in the build
container (needs /root/reference):

    python oracle/make_golden_trainer.py        # -> tests/golden/pretrain_step.npz, tests/golden/reconstructor.npz

a19: `training/pretrain_trainer.py:OpenESSPretrainModel` is imported UNMODIFIED (third-party imports it does not need for the step
-- matplotlib, albumentations, the HuggingFace-shadowed `datasets` namespace, mmcv behind `models/__init__` -- are stubbed in
sys.modules), instantiated with `object.__new__` (its __init__ wants datasets and checkpoints on disk), given the reference's own
module classes with seeded weights, and driven through `createOptimizerDict`, `task_train_step` (:364-372, :427-472),
`trainTaskStepPretrain` (:550-562) and `train_step` (:324-361).  Losses, every gradient and the parameters after one AdamW step
are stored.  The weights come from the goldens already committed (e2vid_tiny.npz, semseg_tiny.npz) and from
tests/seeded_weights.py (the teacher, 24 M parameters: too large to commit, deterministic from its seed).
a9: `e2vid/image_reconstructor.py:ImageReconstructor.update_reconstruction` (:80-123) with `CudaTimer -> Timer` (the CUDA-event
timer cannot run on a CPU-only box), three recurrent steps, all returned tensors stored."""
import logging
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("OPENESS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "..", "tests"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, REF)
    plt = _stub("matplotlib.pyplot", cm=SimpleNamespace(Blues=None))
    _stub("matplotlib", pyplot=plt)
    _stub("albumentations")
    ds = _stub("datasets")
    ds.__path__ = [os.path.join(REF, "datasets")]
    _stub("datasets.wrapper_dataloader", WrapperDataset=object)
    models = _stub("models")                               # models/__init__.py star-imports maskclip_model (needs mmcv)
    models.__path__ = [os.path.join(REF, "models")]
    from models.image_model import DilationFeatureExtractor, Preprocessing
    models.Preprocessing = Preprocessing
    models.maskClipFeatureExtractor = object               # constructed by other trainers, never by this one
    import e2vid.image_reconstructor as ir
    import e2vid.utils.inference_utils as iu
    from e2vid.utils.timers import Timer
    ir.CudaTimer = Timer
    iu.CudaTimer = Timer
    import training.pretrain_trainer as pt                 # the reference trainer module, unmodified
    from e2vid.model.model import E2VIDRecurrent
    from models.style_networks import SemSegE2VID
    from utils.loss_functions import NCELoss, TaskLoss
    return SimpleNamespace(pt=pt, ir=ir, E2VIDRecurrent=E2VIDRecurrent, SemSegE2VID=SemSegE2VID,
                           DilationFeatureExtractor=DilationFeatureExtractor, NCELoss=NCELoss, TaskLoss=TaskLoss)



def reference_options():
    """The argparse defaults config/settings.py:44-49 builds (e2vid/options/inference_options.py), use_gpu off."""
    import argparse
    from e2vid.options.inference_options import set_inference_options
    parser = argparse.ArgumentParser()
    set_inference_options(parser)
    opts, _ = parser.parse_known_args([])
    opts.use_gpu = False
    return opts


def _cfg(z):
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = (v == "True") if str(v) in ("True", "False") else (int(v) if str(v).isdigit() else str(v))
    return cfg


def reference_modules(R):
    from seeded_weights import seeded_state_dict
    z = np.load(os.path.join(OUT, "e2vid_tiny.npz"))
    e2vid = R.E2VIDRecurrent(_cfg(z))
    e2vid.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}, strict=True)
    zs = np.load(os.path.join(OUT, "semseg_tiny.npz"))
    K = int(zs["K"])
    back = R.SemSegE2VID(input_c=32, output_c=K, skip_connect=True, skip_type='concat', text_embeddings_path=None)
    back.load_state_dict({k[4:]: torch.from_numpy(zs[k]) for k in zs.files if k.startswith("sd__")}, strict=True)
    teacher = R.DilationFeatureExtractor(image_weights=None)
    teacher.load_state_dict(seeded_state_dict(teacher, 77), strict=True)
    return e2vid, back, teacher, K


def step_inputs(K, Bn=2, H=32, W=48, S=10, steps=3):
    g = torch.Generator().manual_seed(11)
    event = torch.randn(Bn, 5 * steps, H, W, generator=g)
    event[torch.rand(event.shape, generator=g) < 0.6] = 0
    frame = torch.rand(Bn, 3, H, W, generator=g)
    pl = torch.randint(0, K, (Bn, H, W), generator=g)
    pl[torch.rand(pl.shape, generator=g) < 0.03] = 255
    sp = torch.randint(0, S, (Bn, H, W), generator=g)
    return event, frame, pl, sp, S, steps


STRIDE = 7


def golden_pretrain_step(R):
    torch.manual_seed(1205)
    e2vid, back, teacher, K = reference_modules(R)
    event, frame, pl, sp, S, steps = step_inputs(K)
    H, W = event.shape[-2:]
    dev = torch.device("cpu")
    opts = reference_options()
    logger = logging.getLogger("golden")
    settings = SimpleNamespace(config_option='frame2voxel', unfrozen_e2vid=False, if_spatial_contrastive=True,
                               if_dense_clip_supervision=True, nr_events_data_b=steps, input_channels_b=5,
                               superpixel_size=S, weight_task_loss=1.0, if_switchable_train=False, use_amp=False,
                               lr_voxel=1e-3, lr_frame=1e-3, logger=logger, task_loss=['dice', 'cross_entropy'],
                               semseg_num_classes=K, semseg_ignore_label=255)
    T = R.pt.OpenESSPretrainModel
    tr = object.__new__(T)
    tr.is_training, tr.settings, tr.device, tr.epoch_count = True, settings, dev, 0
    for p in e2vid.parameters():                                            # pretrain_trainer.py:154-157
        p.requires_grad = False
    e2vid.eval()
    tr.front_end_sensor_b, tr.task_backend, tr.model_frame = e2vid, back, teacher
    tr.models_dict = {"front_sensor_b": e2vid, "back_end": back, "model_frame": teacher}
    tr.reconstructor = R.ir.ImageReconstructor(e2vid, H, W, 5, dev, opts)
    tr.task_loss = R.TaskLoss(losses=settings.task_loss, gamma=2.0, num_classes=K, ignore_index=255, reduction='mean')
    tr.nce_loss = R.NCELoss(temperature=0.07)
    T.createOptimizerDict(tr)                                               # :211-274 (two AdamW)

    sd_back = {k: v.clone() for k, v in back.state_dict().items()}
    sd_teacher = {k: v.clone() for k, v in teacher.state_dict().items()}
    batch = (event, None, frame, pl, sp)
    t_loss, losses, outputs = T.task_train_step(tr, batch)                  # :364-372, :427-472
    t_loss.backward()
    d = {"event": event.numpy(), "frame": frame.numpy(), "pl": pl.numpy(), "sp": sp.numpy(), "S": np.array(S),
         "steps": np.array(steps), "K": np.array(K), "total": np.array(t_loss.item()),
         "nce": np.array(losses["contrastive_nce_loss"].item()), "dense": np.array(losses["dense_clip_loss"].item()),
         "logits": outputs["pred"][1].detach().numpy() if "pred" in outputs else np.zeros(0, np.float32)}
    grads = {}
    for prefix, m in (("back_end.", back), ("model_frame.", teacher)):
        for n, p in m.named_parameters():
            if p.grad is not None:
                grads[prefix + n] = p.grad.numpy().copy()
    d["nograd"] = np.array([prefix + n for prefix, m in (("back_end.", back), ("model_frame.", teacher))
                            for n, p in m.named_parameters() if p.grad is None])
    for n, v in grads.items():                          # the 2048 x 256 teacher decoder weight is stored strided (file size)
        d["grad__" + n] = v if v.size < 100_000 else v.reshape(-1)[::STRIDE].copy()
    # the full optimiser step from the same starting point: train_step = zero_grad + task_train_step + backward + 2 x AdamW.step
    back.load_state_dict(sd_back)
    teacher.load_state_dict(sd_teacher)
    for m in (back, teacher):
        for p in m.parameters():
            p.grad = None
    T.createOptimizerDict(tr)
    _, _, final_loss = T.train_step(tr, batch)                              # :324-361
    d["step_total"] = np.array(final_loss.item())
    for n in ("decoder_ch256.0.weight", "decoder_ch256.0.bias", "decoder_scale_1.0.model.1.weight", "text_embeddings"):
        t = dict(back.named_parameters()).get(n)
        if t is not None:
            d["after__back_end." + n] = t.detach().numpy().copy()
    for n, p in teacher.named_parameters():
        if n.startswith("decoder"):
            v = p.detach().numpy()
            d["after__model_frame." + n] = v.copy() if v.size < 100_000 else v.reshape(-1)[::STRIDE].copy()
    d["stride"] = np.array(STRIDE)
    np.savez_compressed(os.path.join(OUT, "pretrain_step.npz"), **d)
    print("pretrain_step.npz", os.path.getsize(os.path.join(OUT, "pretrain_step.npz")) // 1024, "KiB",
          {k: float(d[k]) for k in ("total", "nce", "dense", "step_total")}, len(grads), "gradients")


def golden_reconstructor(R):
    """ImageReconstructor.update_reconstruction (a9): three steps, tiny E2VID, H x W not a multiple of 8 (reflection pad + crop)."""
    e2vid, _, _, _ = reference_modules(R)
    e2vid.eval()
    H, W = 30, 44
    dev = torch.device("cpu")
    opts = reference_options()
    rec = R.ir.ImageReconstructor(e2vid, H, W, 5, dev, opts)
    rng = np.random.default_rng(99)
    d = {"H": np.array(H), "W": np.array(W)}
    for i in range(3):
        ev = rng.normal(0, 1, (2, 5, H, W)).astype(np.float32)
        ev[rng.random(ev.shape) < 0.6] = 0
        img, states, latent = rec.update_reconstruction(torch.from_numpy(ev.copy()))
        d[f"in{i}"] = ev
        d[f"img{i}"] = img.detach().numpy()
        for kk, vv in latent.items():
            d[f"latent{i}__{kk}"] = vv.detach().numpy()
        for li, (h, c) in enumerate(states):
            d[f"state{i}__{li}__h"], d[f"state{i}__{li}__c"] = h.detach().numpy(), c.detach().numpy()
    # an all-zero event tensor takes the `num_nonzeros > 0` false branch of EventPreprocessor (inference_utils.py:81)
    rec.last_states_for_each_channel = {'grayscale': None}
    img, _, latent = rec.update_reconstruction(torch.zeros(1, 5, H, W))
    d["zero_img"] = img.detach().numpy()
    d["zero_latent8"] = latent[8].detach().numpy()
    # `standardization=True` branch (:108-113) and the post-processing of the reconstruction that run_reconstruction.py
    # applies before writing PNGs (image_reconstructor.py:126-140: UnsharpMaskFilter + IntensityRescaler, inference_utils.py:90-129, 234-252)
    rec2 = R.ir.ImageReconstructor(e2vid, H, W, 5, dev, opts, standardization=True)
    img_s, _, _ = rec2.update_reconstruction(torch.from_numpy(d["in0"].copy()))
    d["std_img0"] = img_s.detach().numpy()
    post = R.ir.PostProcessor(dev, opts)
    d["post_img0"] = post.process(torch.from_numpy(d["img0"].copy())).numpy()
    opts_hdr = reference_options()
    opts_hdr.auto_hdr = True
    post_hdr = R.ir.PostProcessor(dev, opts_hdr)
    for i in range(3):
        d[f"post_hdr_img{i}"] = post_hdr.process(torch.from_numpy(d[f"img{i}"].copy())).numpy()
    d["unsharp_amount"], d["unsharp_sigma"] = np.array(opts.unsharp_mask_amount), np.array(opts.unsharp_mask_sigma)
    np.savez_compressed(os.path.join(OUT, "reconstructor.npz"), **d)
    print("reconstructor.npz", os.path.getsize(os.path.join(OUT, "reconstructor.npz")) // 1024, "KiB")


def golden_openess_step(R):
    """training/openess_trainer.py:OpenESSModel, `frame2recon` branch (the only coherent one):
    task_train_step (:360-372, 476-529) and train_step (:339-358) of the unmodified class with two reference
    deeplabv3_resnet50 networks (seeded weights, Dropout p = 0 for determinism)."""
    from seeded_weights import seeded_state_dict
    import training.openess_trainer as ot
    from models.deeplabv3 import deeplabv3_resnet50
    torch.manual_seed(1205)
    K = 11
    nets = []
    for seed in (4, 5):
        m = deeplabv3_resnet50(num_classes=K, text_embeddings_path=None, output_stride=32, pretrained_backbone='')
        m.load_state_dict(seeded_state_dict(m, seed), strict=True)
        m.classifier.ASPP.project[3].p = 0.0
        nets.append(m)
    model_frame, model_recon = nets
    rng = np.random.default_rng(21)
    B, H, W = 2, 64, 96
    frame = torch.from_numpy(rng.random((B, 3, H, W)).astype(np.float32))
    recon = torch.from_numpy(rng.random((B, 3, H, W)).astype(np.float32))
    pl = rng.integers(0, K, (B, H, W))
    pl[rng.random(pl.shape) < 0.03] = 255
    pl = torch.from_numpy(pl.astype(np.int64))
    sp = torch.from_numpy(rng.integers(0, 30, (B, H, W)).astype(np.int64))
    settings = SimpleNamespace(config_option='frame2recon', if_spatial_contrastive=True, weight_task_loss=1.0,
                               lr_recon=1e-3, lr_frame=1e-3, task_loss=['dice', 'cross_entropy'], semseg_num_classes=K,
                               semseg_ignore_label=255, logger=logging.getLogger("golden"))
    T = ot.OpenESSModel
    tr = object.__new__(T)
    tr.is_training, tr.settings, tr.device, tr.epoch_count = True, settings, torch.device("cpu"), 0
    tr.model_frame, tr.model_recon = model_frame, model_recon
    tr.models_dict = {"model_recon": model_recon, "model_frame": model_frame}
    tr.task_loss = R.TaskLoss(losses=settings.task_loss, gamma=2.0, num_classes=K, ignore_index=255, reduction='mean')
    tr.l1_loss = torch.nn.L1Loss()
    tr.nce_loss = R.NCELoss(temperature=0.07)
    T.createOptimizerDict(tr)
    sd = [{k: v.clone() for k, v in m.state_dict().items()} for m in nets]
    batch = (frame, None, recon, pl, sp)
    t_loss, losses, _ = T.task_train_step(tr, batch)
    t_loss.backward()
    d = {"frame": frame.numpy(), "recon": recon.numpy(), "pl": pl.numpy(), "sp": sp.numpy(), "K": np.array(K),
         "total": np.array(t_loss.item())}
    for k, v in losses.items():
        d["loss__" + k] = np.array(v.item())
    picks = ("classifier.text_embeddings", "classifier.ASPP.project.0.weight", "backbone.conv1.weight",
             "backbone.layer4.2.conv3.weight", "classifier.classifier.0.weight")
    for prefix, m in (("model_frame.", model_frame), ("model_recon.", model_recon)):
        named = dict(m.named_parameters())
        for n in picks:
            g = named[n].grad.numpy()
            d["grad__" + prefix + n] = g if g.size < 50_000 else g.reshape(-1)[::STRIDE * 7].copy()
        d["nograd__" + prefix] = np.array([n for n, p in named.items() if p.grad is None])
    for m, s0 in zip(nets, sd):
        m.load_state_dict(s0)
        for p in m.parameters():
            p.grad = None
    T.createOptimizerDict(tr)
    _, _, final = T.train_step(tr, batch)
    d["step_total"] = np.array(final.item())
    for prefix, m in (("model_frame.", model_frame), ("model_recon.", model_recon)):
        named = dict(m.named_parameters())
        for n in ("classifier.text_embeddings", "backbone.conv1.weight"):
            v = named[n].detach().numpy()
            d["after__" + prefix + n] = v.copy() if v.size < 50_000 else v.reshape(-1)[::STRIDE * 7].copy()
    d["stride"] = np.array(STRIDE * 7)
    np.savez_compressed(os.path.join(OUT, "openess_step.npz"), **d)
    print("openess_step.npz", os.path.getsize(os.path.join(OUT, "openess_step.npz")) // 1024, "KiB",
          {k: float(v) for k, v in d.items() if k.startswith("loss__") or k in ("total", "step_total")})


def main():
    torch.set_num_threads(1)
    R = import_reference()
    if "--openess-only" not in sys.argv:
        golden_reconstructor(R)
        golden_pretrain_step(R)
    golden_openess_step(R)


if __name__ == "__main__":
    main()
