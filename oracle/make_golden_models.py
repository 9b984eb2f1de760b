"""Golden vectors for the model-side rows (a9/a10 E2VID recurrent encoder, a18 consistency losses), produced by
the REFERENCE's own modules on CPU in the build container (needs /root/reference):

    python oracle/make_golden_models.py        # -> tests/golden/e2vid_tiny.npz, tests/golden/consistency.npz

e2vid/model/model.py:E2VIDRecurrent is imported unmodified (sys.modules is pre-seeded for the `e2vid` namespace
package; `e2vid.base` needs nothing that is missing here).  A tiny config keeps the weights small enough to commit."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("OPENESS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CFG = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
       'base_num_channels': 4, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, REF)
    from e2vid.model.model import E2VIDRecurrent          # the reference class, unmodified
    torch.manual_seed(1205)
    m = E2VIDRecurrent(CFG).eval()
    with torch.no_grad():                                  # non-trivial BN statistics / affine
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.2)
    rng = np.random.default_rng(1205)
    steps = [rng.normal(0, 1, (2, 5, 24, 32)).astype(np.float32) for _ in range(3)]
    for s in steps:
        s[rng.random(s.shape) < 0.6] = 0
    out = {"cfg_keys": np.array(list(CFG.keys())), "cfg_vals": np.array([str(v) for v in CFG.values()])}
    for k, v in m.state_dict().items():
        out["sd__" + k] = v.numpy()
    states = None
    with torch.no_grad():
        for i, s in enumerate(steps):
            img, states, latent = m(torch.from_numpy(s), states)
            out[f"in{i}"] = s
            out[f"img{i}"] = img.numpy()
            for kk, vv in latent.items():
                out[f"latent{i}__{kk}"] = vv.numpy()
            for li, (h, c) in enumerate(states):
                out[f"state{i}__{li}__h"], out[f"state{i}__{li}__c"] = h.numpy(), c.numpy()
    np.savez_compressed(os.path.join(OUT, "e2vid_tiny.npz"), **out)

    # a18: torch.nn.L1Loss and mean(1 - cosine_similarity(dim=1)) with gradients (openess_trainer.py:456-462)
    a = torch.from_numpy(rng.normal(0, 1, (2, 16, 14, 18)).astype(np.float32)).requires_grad_(True)
    b = torch.from_numpy(rng.normal(0, 1, (2, 16, 14, 18)).astype(np.float32)).requires_grad_(True)
    l1 = torch.nn.L1Loss()(a, b)
    l1.backward()
    la = torch.from_numpy(rng.normal(0, 2, (2, 11, 14, 18)).astype(np.float32)).requires_grad_(True)
    lb = torch.from_numpy(rng.normal(0, 2, (2, 11, 14, 18)).astype(np.float32)).requires_grad_(True)
    cs = torch.mean(1 - torch.nn.functional.cosine_similarity(la, lb, dim=1))
    cs.backward()
    np.savez_compressed(os.path.join(OUT, "consistency.npz"), a=a.detach().numpy(), b=b.detach().numpy(),
                        l1=np.array(l1.item()), da=a.grad.numpy(), db=b.grad.numpy(), la=la.detach().numpy(),
                        lb=lb.detach().numpy(), cos=np.array(cs.item()), dla=la.grad.numpy(), dlb=lb.grad.numpy())
    for f in ("e2vid_tiny.npz", "consistency.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
    golden_semseg()




def golden_semseg():
    """models/style_networks.py:SemSegE2VID (reference class, tiny width) forward + gradients + superpixel pooling."""
    sys.modules.setdefault("models", types.ModuleType("models")).__path__ = [os.path.join(REF, "models")]
    from models.style_networks import SemSegE2VID          # the reference class, unmodified
    torch.manual_seed(7)
    K, S, B, H, W = 6, 10, 2, 16, 24
    m = SemSegE2VID(input_c=32, output_c=K, skip_connect=True, skip_type='concat', text_embeddings_path=None)
    with torch.no_grad():
        m.text_embeddings.normal_(0, 0.3)
        for p in m.parameters():
            if p.ndim == 4:
                p.mul_(8.0)                                 # gaussian_weights_init std 0.02 is too small to exercise anything
    rng = np.random.default_rng(7)
    lat = {8: torch.from_numpy(rng.normal(0, 1, (B, 32, H // 8, W // 8)).astype(np.float32)).requires_grad_(True),
           4: torch.from_numpy(rng.normal(0, 1, (B, 16, H // 4, W // 4)).astype(np.float32)).requires_grad_(True),
           2: torch.from_numpy(rng.normal(0, 1, (B, 8, H // 2, W // 2)).astype(np.float32)).requires_grad_(True),
           1: torch.from_numpy(rng.normal(0, 1, (B, 4, H, W)).astype(np.float32))}
    sp = torch.from_numpy(rng.integers(0, S, (B, H, W)).astype(np.int64))
    out, x256 = m(lat)
    superpixels = torch.arange(0, B * S, S)[:, None, None] + sp      # pretrain_trainer.py:446-459
    sI = superpixels.flatten()
    with torch.no_grad():
        oh = torch.sparse_coo_tensor(torch.stack((sI, torch.arange(sI.shape[0])), 0), torch.ones(sI.shape[0]))
    k = oh @ x256.permute(0, 2, 3, 1).flatten(0, 2)
    k = k / (torch.sparse.sum(oh, 1).to_dense()[:, None] + 1e-6)
    loss = out[1].square().mean() + 3.0 * k.square().mean()
    loss.backward()
    d = {"K": np.array(K), "S": np.array(S), "sp": sp.numpy(), "logits": out[1].detach().numpy(),
         "out2": out[2].detach().numpy(), "out4": out[4].detach().numpy(), "x256": x256.detach().numpy(),
         "k": k.detach().numpy(), "loss": np.array(loss.item())}
    for kk, v in lat.items():
        d[f"lat{kk}"] = v.detach().numpy()
        if v.grad is not None:
            d[f"dlat{kk}"] = v.grad.numpy()
    for n, p in m.state_dict().items():
        d["sd__" + n] = p.numpy()
    for n, p in m.named_parameters():
        if p.grad is not None and n.startswith(("decoder_ch", "text_emb", "decoder_scale_4")):
            d["grad__" + n] = p.grad.numpy()
    d["nograd"] = np.array([n for n, p in m.named_parameters() if p.grad is None])
    np.savez_compressed(os.path.join(OUT, "semseg_tiny.npz"), **d)
    print("semseg_tiny.npz", os.path.getsize(os.path.join(OUT, "semseg_tiny.npz")) // 1024, "KiB")


CFG_TC = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
          'base_num_channels': 32, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}


def golden_e2vid_full_width():
    """The real E2VID-lightweight width (ConvLSTM hidden sizes 64 / 128 / 256: the sizes the tcgen05 ConvLSTM kernel
    serves) at a small spatial size.  10.7 M weights are not committed: tests/seeded_weights.py regenerates them."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from seeded_weights import seeded_state_dict
    from e2vid.model.model import E2VIDRecurrent          # the reference class, unmodified
    m = E2VIDRecurrent(CFG_TC).eval()
    m.load_state_dict(seeded_state_dict(m, 1205), strict=True)
    rng = np.random.default_rng(99)
    steps = [rng.normal(0, 1, (1, 5, 24, 40)).astype(np.float32) for _ in range(3)]
    for s_ in steps:
        s_[rng.random(s_.shape) < 0.6] = 0
    out = {"cfg_keys": np.array(list(CFG_TC.keys())), "cfg_vals": np.array([str(v) for v in CFG_TC.values()]),
           "seed": np.array(1205)}
    states = None
    with torch.no_grad():
        for i, s_ in enumerate(steps):
            img, states, latent = m(torch.from_numpy(s_), states)
            out[f"in{i}"] = s_
            out[f"img{i}"] = img.numpy()                   # reconstruction of every step (resblocks + decoders + pred + sigmoid)
    for kk, vv in latent.items():                          # outputs of the LAST step (they depend on all three)
        out[f"latent__{kk}"] = vv.numpy()
    for li, (h, c) in enumerate(states):
        out[f"state__{li}__c"] = c.numpy()
    np.savez_compressed(os.path.join(OUT, "e2vid_full_width.npz"), **out)
    print("e2vid_full_width.npz", os.path.getsize(os.path.join(OUT, "e2vid_full_width.npz")) // 1024, "KiB")


def golden_teacher():
    """models/image_model.py:DilationFeatureExtractor (reference class, unmodified, image_weights=None) in TRAIN mode --
    how the trainers run it (pretrain_trainer.py:370-371).  24 M weights regenerated from tests/seeded_weights.py."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from seeded_weights import seeded_state_dict
    sys.modules.setdefault("models", types.ModuleType("models")).__path__ = [os.path.join(REF, "models")]
    from models.image_model import DilationFeatureExtractor
    m = DilationFeatureExtractor(image_weights=None)
    m.load_state_dict(seeded_state_dict(m, 77), strict=True)
    m.train()
    rng = np.random.default_rng(5)
    x = rng.random((2, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        feats = m.encoder(torch.from_numpy(x))
    m.load_state_dict(seeded_state_dict(m, 77), strict=True)      # undo the running-stat update of the probe call
    y = m(torch.from_numpy(x))
    sd = m.state_dict()
    out = {"seed": np.array(77), "x": x, "feats_sub": feats[:, ::16].numpy(), "y_sub": y.detach()[:, ::8, ::4, ::4].numpy(),
           "feats_absmax": np.array(float(feats.abs().max())),
           "rm_l4": sd["encoder.layer4.2.bn3.running_mean"].numpy(), "rv_l1": sd["encoder.layer1.0.bn1.running_var"].numpy(),
           "rv_ds": sd["encoder.layer2.0.downsample.1.running_var"].numpy(),
           "nbt": sd["encoder.layer3.5.bn2.num_batches_tracked"].numpy()}
    np.savez_compressed(os.path.join(OUT, "teacher_r50.npz"), **out)
    print("teacher_r50.npz", os.path.getsize(os.path.join(OUT, "teacher_r50.npz")) // 1024, "KiB")


def golden_deeplab():
    """models/deeplabv3.py:deeplabv3_resnet50 (reference class, unmodified) as BASELINE config 4 builds it: K = 11,
    yaml output_stride 32 (-> effective 16), if_finetuning + frozen_backbone.  Train-mode forward (batch-stat BN, Dropout
    disabled for determinism by p = 0) + head gradients, and eval-mode forward.  41 M weights from tests/seeded_weights.py."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from seeded_weights import seeded_state_dict
    sys.modules.setdefault("models", types.ModuleType("models")).__path__ = [os.path.join(REF, "models")]
    from models.deeplabv3 import deeplabv3_resnet50
    m = deeplabv3_resnet50(num_classes=11, text_embeddings_path=None, output_stride=32, pretrained_backbone='',
                           if_finetuning=True, frozen_backbone=True)
    sd0 = seeded_state_dict(m, 4)
    m.load_state_dict(sd0, strict=True)
    m.classifier.ASPP.project[3].p = 0.0                   # Dropout(0.1) is random: disabled in the golden AND the test
    rng = np.random.default_rng(8)
    x = rng.random((2, 3, 64, 96)).astype(np.float32)
    out = {"seed": np.array(4), "x": x, "nkeys": np.array(len(sd0)), "nparams": np.array(sum(p.numel() for p in m.parameters()))}
    m.eval()
    with torch.no_grad():
        le, fe = m(torch.from_numpy(x))
    out["eval_logits_sub"], out["eval_feats_sub"] = le[:, :, ::2, ::2].numpy(), fe[:, ::8, ::4, ::4].numpy()
    m.train()
    lt, ft = m(torch.from_numpy(x))
    (lt.square().mean() + ft.square().mean()).backward()
    out["train_logits_sub"], out["train_feats_sub"] = lt.detach()[:, :, ::2, ::2].numpy(), ft.detach()[:, ::8, ::4, ::4].numpy()
    out["grad_text"] = m.classifier.text_embeddings.grad.numpy()
    out["grad_proj_sub"] = m.classifier.ASPP.project[0].weight.grad[::4, ::16, 0, 0].numpy()
    out["rm_l4"] = m.state_dict()["backbone.layer4.2.bn3.running_mean"].numpy()
    out["backbone_has_grad"] = np.array(any(p.grad is not None for p in m.backbone.parameters()))
    np.savez_compressed(os.path.join(OUT, "deeplab_r50.npz"), **out)
    print("deeplab_r50.npz", os.path.getsize(os.path.join(OUT, "deeplab_r50.npz")) // 1024, "KiB")


def golden_resnet18():
    """models/_resnet.py:resnet18 (reference function, unmodified, pretrained='') -- row a14': eval-mode and train-mode
    forward (logits + layer4 map + running statistics).  11.7 M weights regenerated from tests/seeded_weights.py."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from seeded_weights import seeded_state_dict
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_resnet", os.path.join(REF, "models", "_resnet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m = mod.resnet18(pretrained='')
    sd0 = seeded_state_dict(m, 18)
    m.load_state_dict(sd0, strict=True)
    rng = np.random.default_rng(18)
    x = rng.random((2, 3, 96, 160)).astype(np.float32)

    def feats(net, t):
        t = net.maxpool(net.relu(net.bn1(net.conv1(t))))
        return net.layer4(net.layer3(net.layer2(net.layer1(t))))

    out = {"seed": np.array(18), "x": x, "nkeys": np.array(len(sd0)), "nparams": np.array(sum(p.numel() for p in m.parameters()))}
    m.eval()
    with torch.no_grad():
        out["eval_logits"] = m(torch.from_numpy(x)).numpy()
        out["eval_feats"] = feats(m, torch.from_numpy(x)).numpy()
    m.train()
    with torch.no_grad():
        out["train_logits"] = m(torch.from_numpy(x)).numpy()
    sd = m.state_dict()
    out["rm_l4"] = sd["layer4.1.bn2.running_mean"].numpy()
    out["rv_stem"] = sd["bn1.running_var"].numpy()
    out["nbt"] = sd["layer2.0.downsample.1.num_batches_tracked"].numpy()
    np.savez_compressed(os.path.join(OUT, "resnet18.npz"), **out)
    print("resnet18.npz", os.path.getsize(os.path.join(OUT, "resnet18.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if "--resnet18" in sys.argv:
        torch.set_num_threads(4)
        golden_resnet18()
    elif "--deeplab" in sys.argv:
        sys.path.insert(0, REF)
        torch.set_num_threads(4)
        golden_deeplab()
    elif "--teacher" in sys.argv:
        sys.path.insert(0, REF)
        torch.set_num_threads(4)
        golden_teacher()
    elif "--e2vid-full" in sys.argv:
        sys.path.insert(0, REF)
        torch.set_num_threads(1)
        golden_e2vid_full_width()
    elif "--semseg" in sys.argv:
        sys.path.insert(0, REF)
        torch.set_num_threads(1)
        golden_semseg()
    else:
        main()
