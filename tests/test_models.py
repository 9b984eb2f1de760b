"""E2VID recurrent encoder mirror (a9/a10) and consistency losses (a18) against goldens produced by the
reference's own modules (oracle/make_golden_models.py).  Tolerances: float32 convolutions in a different
summation order (cuDNN vs the reference's CPU kernels, folded BN): 2e-4 abs on O(1) activations."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def _build(z, latent_only):
    from openess_b200.e2vid.model.model import E2VIDRecurrent
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = (v == "True") if str(v) in ("True", "False") else (int(v) if str(v).isdigit() else str(v))
    m = E2VIDRecurrent(cfg, latent_only=latent_only)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    m.load_state_dict(sd, strict=True)          # same module tree + parameter names as the reference checkpoint
    return m.eval()


def test_e2vid_mirror_state_dict_and_cpu_forward():
    z = load_golden("e2vid_tiny")
    m = _build(z, latent_only=False)
    states = None
    with torch.no_grad():
        for i in range(3):
            img, states, latent = m(torch.from_numpy(z[f"in{i}"]), states)
            np.testing.assert_allclose(img.numpy(), z[f"img{i}"], atol=2e-5)
            assert sorted(latent) == [1, 2, 4, 8]
            for k in latent:
                np.testing.assert_allclose(latent[k].numpy(), z[f"latent{i}__{k}"], atol=2e-5)
    assert sum(p.numel() for p in m.parameters()) == sum(z[k].size for k in z.files if k.startswith("sd__") and
                                                         "running" not in k and "num_batches" not in k)


@pytest.mark.gpu
def test_e2vid_latent_only_gpu_fused_gates_and_folded_bn():
    z = load_golden("e2vid_tiny")
    dev = torch.device("cuda:0")
    m = _build(z, latent_only=True).to(dev).fold_bn()
    states = None
    with torch.no_grad():
        for i in range(3):
            img, states, latent = m(torch.from_numpy(z[f"in{i}"]).to(dev), states)
            assert img is None                                 # decoders / pred / sigmoid skipped
            for k in latent:
                np.testing.assert_allclose(latent[k].cpu().numpy(), z[f"latent{i}__{k}"], atol=2e-4)
            for li, (h, c) in enumerate(states):
                np.testing.assert_allclose(h.cpu().numpy(), z[f"state{i}__{li}__h"], atol=2e-4)
                np.testing.assert_allclose(c.cpu().numpy(), z[f"state{i}__{li}__c"], atol=2e-4)


@pytest.mark.gpu
def test_e2vid_full_width_tensor_core_convlstm_vs_reference_golden():
    """Real E2VID-lightweight width (ConvLSTM hidden 64 / 128 / 256): every ConvLSTM step runs on the tcgen05 kernel.
    Golden = the REFERENCE E2VIDRecurrent on CPU fp32 (oracle/make_golden_models.py --e2vid-full), weights regenerated
    from the committed seed.  Tolerances: strict-fp32 path (cuDNN + fused gates) 3e-4; tensor-core path (TF32 operands,
    what torch's default cuDNN conv does on the reference's GPU run) 1e-2 abs on O(1) activations after three recurrent
    steps through three levels."""
    from seeded_weights import seeded_state_dict
    from openess_b200 import _lib
    from openess_b200.e2vid.model import model as mm
    z = load_golden("e2vid_full_width")
    dev = torch.device("cuda:0")
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = (v == "True") if str(v) in ("True", "False") else (int(v) if str(v).isdigit() else str(v))
    m = mm.E2VIDRecurrent(cfg, latent_only=True)
    m.load_state_dict(seeded_state_dict(m, int(z["seed"])), strict=True)
    m = m.eval().to(dev).fold_bn()
    errs = {}
    bf16_was = mm.CONVLSTM_BF16
    for mode in ("fp32", "tf32", "bf16"):
        use_tc = mode != "fp32"
        mm.USE_TENSOR_CORES = use_tc
        mm.CONVLSTM_BF16 = mode == "bf16"
        try:
            states = None
            with _lib.profile() as prof:
                with torch.no_grad():
                    for i in range(3):
                        _, states, latent = m(torch.from_numpy(z[f"in{i}"]).to(dev), states)
            # stated tolerances on O(1) activations after three recurrent steps through three levels: strict fp32 3e-4, TF32
            # operands 1e-2, bf16 ConvLSTM operands (default for the frozen encoder) 3e-2
            tol = {"fp32": 3e-4, "tf32": 1e-2, "bf16": 3e-2}[mode]
            worst = 0.0
            for k in latent:
                worst = max(worst, float(np.abs(latent[k].cpu().numpy() - z[f"latent__{k}"]).max()))
                np.testing.assert_allclose(latent[k].cpu().numpy(), z[f"latent__{k}"], atol=tol)
            for li, (h, c) in enumerate(states):
                np.testing.assert_allclose(c.cpu().numpy(), z[f"state__{li}__c"], atol=tol)
            errs[mode] = worst
            # 3 levels x 3 steps of ConvLSTM and encoder convs + 3 head convs on the tensor cores, none otherwise
            assert prof.kernels.get("tc_convlstm_step", (0, 0.0))[0] == (9 if mode == "tf32" else 0)
            assert prof.kernels.get("tc_convlstm_step_bf16", (0, 0.0))[0] == (9 if mode == "bf16" else 0)
            # bf16 mode: the level-2 / level-3 encoder convs read the previous level's bf16 hidden state (kind::f16 kernel)
            nconv = prof.kernels.get("tc_conv2d", (0, 0.0))[0] + prof.kernels.get("tc_conv2d_bf16", (0, 0.0))[0]
            assert nconv == (12 if use_tc else 0)
            assert prof.kernels.get("tc_conv2d_bf16", (0, 0.0))[0] == (6 if mode == "bf16" else 0)
        finally:
            mm.USE_TENSOR_CORES = True
            mm.CONVLSTM_BF16 = bf16_was
    print("max |latent error| fp32 path %.2e, TF32 tensor-core path %.2e, bf16 ConvLSTM path %.2e" % (errs["fp32"], errs["tf32"], errs["bf16"]))


@pytest.mark.gpu
def test_e2vid_online_reconstruction_on_own_kernels_vs_reference_golden():
    """the image branch (resblocks, transposed-conv decoders with fused skip sums, prediction layer +
    sigmoid) on hand-written kernels.  Golden = the reference E2VIDRecurrent's `img` of every step (CPU fp32).  Tolerance:
    5e-3 on a sigmoid output in (0, 1) after TF32 convs (measured value printed); the torch formulation of the same mirror
    on the GPU (strict fp32) 1e-4."""
    from seeded_weights import seeded_state_dict
    from openess_b200 import _lib, ops
    from openess_b200.e2vid.model import model as mm
    z = load_golden("e2vid_full_width")
    dev = torch.device("cuda:0")
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = (v == "True") if str(v) in ("True", "False") else (int(v) if str(v).isdigit() else str(v))
    m = mm.E2VIDRecurrent(cfg, latent_only=False)
    m.load_state_dict(seeded_state_dict(m, int(z["seed"])), strict=True)
    m = m.eval().to(dev).fold_bn()
    errs = {}
    for use_tc in (False, True):
        mm.USE_TENSOR_CORES = use_tc
        try:
            states, worst = None, 0.0
            with _lib.profile() as prof, torch.no_grad():
                for i in range(3):
                    img, states, _ = m(torch.from_numpy(z[f"in{i}"]).to(dev), states)
                    assert img.shape == (1, 1, 24, 40)
                    worst = max(worst, float(np.abs(img.cpu().numpy() - z[f"img{i}"]).max()))
            errs[use_tc] = worst
            if use_tc:      # per step: 2 resblocks x 2 convs + 3 decoders (+ 1 head + 3 encoder convs) on the conv kernel
                nconv = prof.kernels["tc_conv2d"][0] + prof.kernels.get("tc_conv2d_bf16", (0, 0.0))[0]
                assert nconv == 3 * (4 + 4 + 3) and prof.kernels["zero_insert2x_nhwc"][0] == 9
                assert prof.kernels["pred_sigmoid_nhwc"][0] == 3
        finally:
            mm.USE_TENSOR_CORES = True
    print("reconstruction max |err| vs reference: torch fp32 path %.2e, own kernels %.2e" % (errs[False], errs[True]))
    assert errs[False] < 1e-4 and errs[True] < 5e-3
    # the transposed convolution identity on its own: zero insertion + rotated kernel == F.conv_transpose2d
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randint(-3, 4, (2, 32, 5, 7), device=dev, generator=g).float()
    sk = torch.randint(-3, 4, (2, 32, 5, 7), device=dev, generator=g).float()
    w = torch.randint(-2, 3, (32, 16, 5, 5), device=dev, generator=g).float()
    got = ops.conv2d_tc(ops.zero_insert2x_nhwc(x, sk), ops.conv_transpose2x_pack(w), None, 5, 1, 2, 1)
    want = torch.nn.functional.conv_transpose2d(x + sk, w, stride=2, padding=2, output_padding=1)
    assert torch.equal(got, want)                                    # small integers: exact in TF32


@pytest.mark.gpu
def test_convlstm_gates_kernel_vs_torch():
    from openess_b200 import losses
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for shape in [(2, 8, 12, 16), (1, 3, 5, 7)]:               # vectorised and scalar paths
        B, C, H, W = shape
        gates = (torch.randn((B, 4 * C, H, W), generator=g) * 3).to(dev)
        pc = torch.randn(shape, generator=g).to(dev)
        for prev in (pc, None):
            h, c = losses.convlstm_gates(gates, prev)
            i_, r_, o_, c_ = gates.double().chunk(4, 1)
            cell = torch.sigmoid(r_) * (0 if prev is None else prev.double()) + torch.sigmoid(i_) * torch.tanh(c_)
            hid = torch.sigmoid(o_) * torch.tanh(cell)
            torch.testing.assert_close(c.double(), cell, atol=3e-6, rtol=3e-6)
            torch.testing.assert_close(h.double(), hid, atol=3e-6, rtol=3e-6)


@pytest.mark.gpu
def test_image_reconstructor_mirror_runs_recurrence(oracle):
    from types import SimpleNamespace
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    z = load_golden("e2vid_tiny")
    dev = torch.device("cuda:0")
    m = _build(z, latent_only=True).to(dev).fold_bn()
    opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
    rec = ImageReconstructor(m, 24, 32, 5, dev, opts)
    assert rec.crop.is_noop
    rec.last_states_for_each_channel = {'grayscale': None}      # what the trainers do before every sample
    ref_states = None
    for i in range(3):
        x = z[f"in{i}"]
        _, states, latent = rec.update_reconstruction(torch.from_numpy(x))
        with torch.no_grad():                                   # same pipeline composed from checked pieces
            xn, _ = oracle.nonzero_standardize(x)
            _, ref_states, ref_latent = m(torch.from_numpy(xn).to(dev), ref_states)
        for k in latent:
            torch.testing.assert_close(latent[k], ref_latent[k], atol=2e-4, rtol=1e-4)
    assert rec.last_states_for_each_channel['grayscale'] is states
    rec2 = ImageReconstructor(m, 25, 33, 5, dev, opts)          # padding to a multiple of 8 (ReflectionPad2d)
    assert not rec2.crop.is_noop and rec2.crop.width_crop_size == 40 and rec2.crop.height_crop_size == 32
    _, _, lat = rec2.update_reconstruction(torch.zeros(1, 5, 25, 33))
    assert tuple(lat[1].shape[2:]) == (32, 40) and tuple(lat[8].shape[2:]) == (4, 5)


@pytest.mark.gpu
def test_consistency_losses_golden():
    from openess_b200.training.consistency import L1Loss, prediction_consistency
    z = load_golden("consistency")
    dev = torch.device("cuda:0")
    a = torch.from_numpy(z["a"]).to(dev).requires_grad_(True)
    b = torch.from_numpy(z["b"]).to(dev).requires_grad_(True)
    l1 = L1Loss()(a, b)
    assert float(l1.detach()) == pytest.approx(float(z["l1"]), rel=2e-6)
    (2 * l1).backward()
    np.testing.assert_allclose(a.grad.cpu().numpy(), 2 * z["da"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(b.grad.cpu().numpy(), 2 * z["db"], rtol=1e-6, atol=1e-10)
    la = torch.from_numpy(z["la"]).to(dev).requires_grad_(True)
    lb = torch.from_numpy(z["lb"]).to(dev).requires_grad_(True)
    cs = prediction_consistency(la, lb)
    assert float(cs.detach()) == pytest.approx(float(z["cos"]), rel=5e-6)
    cs.backward()
    np.testing.assert_allclose(la.grad.cpu().numpy(), z["dla"], rtol=2e-4, atol=2e-9)
    np.testing.assert_allclose(lb.grad.cpu().numpy(), z["dlb"], rtol=2e-4, atol=2e-9)


@pytest.mark.gpu
def test_image_reconstructor_matches_reference_class_golden():
    """a9 against the REFERENCE's own ImageReconstructor (e2vid/image_reconstructor.py:80-123, run on CPU with CudaTimer -> Timer
    by oracle/make_golden_trainer.py): three recurrent steps at 30 x 44 (reflection-padded to 32 x 48), image, latents, states,
    the all-zero input branch of the EventPreprocessor and `standardization=True`."""
    from types import SimpleNamespace
    from openess_b200.e2vid.image_reconstructor import ImageReconstructor
    z = load_golden("reconstructor")
    ze = load_golden("e2vid_tiny")
    dev = torch.device("cuda:0")
    H, W = int(z["H"]), int(z["W"])
    opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
    m = _build(ze, latent_only=False).to(dev)
    rec = ImageReconstructor(m, H, W, 5, dev, opts)
    assert not rec.crop.is_noop
    for i in range(3):
        img, states, latent = rec.update_reconstruction(torch.from_numpy(z[f"in{i}"]))
        np.testing.assert_allclose(img.cpu().numpy(), z[f"img{i}"], atol=3e-4)
        for k in latent:
            np.testing.assert_allclose(latent[k].cpu().numpy(), z[f"latent{i}__{k}"], atol=3e-4)
        for li, (h, c) in enumerate(states):
            np.testing.assert_allclose(h.cpu().numpy(), z[f"state{i}__{li}__h"], atol=3e-4)
            np.testing.assert_allclose(c.cpu().numpy(), z[f"state{i}__{li}__c"], atol=3e-4)
    assert rec.last_states_for_each_channel['grayscale'] is states
    rec.last_states_for_each_channel = {'grayscale': None}
    img, _, latent = rec.update_reconstruction(torch.zeros(1, 5, H, W))
    np.testing.assert_allclose(img.cpu().numpy(), z["zero_img"], atol=3e-4)
    np.testing.assert_allclose(latent[8].cpu().numpy(), z["zero_latent8"], atol=3e-4)
    rec_s = ImageReconstructor(m, H, W, 5, dev, opts, standardization=True)      # keyword as in the reference signature
    img_s, _, _ = rec_s.update_reconstruction(torch.from_numpy(z["in0"]))
    np.testing.assert_allclose(img_s.cpu().numpy(), z["std_img0"], atol=2e-3)
    # latent-only encoder (what the trainers run): same latents, no image
    m2 = _build(ze, latent_only=True).to(dev).fold_bn()
    rec2 = ImageReconstructor(m2, H, W, 5, dev, opts)
    for i in range(3):
        img, _, latent = rec2.update_reconstruction(torch.from_numpy(z[f"in{i}"]))
        assert img is None
        for k in latent:
            np.testing.assert_allclose(latent[k].cpu().numpy(), z[f"latent{i}__{k}"], atol=3e-4)


@pytest.mark.gpu
def test_post_processor_matches_reference_golden():
    """8f row 3 leftovers: PostProcessor = UnsharpMaskFilter + IntensityRescaler (e2vid/image_reconstructor.py:126-140,
    e2vid/utils/inference_utils.py:90-129, 234-252) as one kernel, against the reference classes' output on the reference's
    own reconstructions; 8-bit quantisation: a pixel may land on the neighbouring level when the float chain differs in the
    last bit, so at most 0.5 % of the pixels may differ, by exactly one level."""
    from types import SimpleNamespace
    from openess_b200.e2vid.image_reconstructor import PostProcessor, gaussian_kernel_5x5
    z = load_golden("reconstructor")
    dev = torch.device("cuda:0")
    opts = SimpleNamespace(unsharp_mask_amount=float(z["unsharp_amount"]), unsharp_mask_sigma=float(z["unsharp_sigma"]),
                           auto_hdr=False, auto_hdr_median_filter_size=10, Imin=0.0, Imax=1.0, bilateral_filter_sigma=0.0)
    k = gaussian_kernel_5x5(1.0)
    assert abs(float(k.sum()) - 1.0) < 1e-6 and tuple(k.shape) == (5, 5)

    def check(got, want):
        d = np.abs(got.cpu().numpy() - want)
        assert float(d.max()) <= 1.0 / 255 + 1e-6
        assert float((d > 1e-6).mean()) < 5e-3

    check(PostProcessor(dev, opts).process(torch.from_numpy(z["img0"])), z["post_img0"])
    opts.auto_hdr = True
    post = PostProcessor(dev, opts)                      # running median of the clipped per-image bounds over three images
    for i in range(3):
        check(post.process(torch.from_numpy(z[f"img{i}"])), z[f"post_hdr_img{i}"])


@pytest.mark.gpu
def test_conv_layer_bn_fold_follows_later_weight_loads():
    """`ConvLayer.fold_bn()` before `load_state_dict` / `.to(device)` must not leave stale folded weights: the fold is
    keyed on its sources and redone lazily."""
    from openess_b200.e2vid.model import model as mm
    g = torch.Generator().manual_seed(2)
    layer = mm.ConvLayer(32, 64, 5, stride=2, padding=2, norm='BN').eval()
    layer.fold_bn()                                          # folded on the CPU, with the initial weights
    sd = {k: torch.randn(v.shape, generator=g) * 0.1 if v.dtype.is_floating_point else v for k, v in layer.state_dict().items()}
    sd["norm_layer.running_var"] = sd["norm_layer.running_var"].abs() + 0.5
    layer.load_state_dict(sd)
    layer = layer.cuda()
    x = torch.randn(2, 32, 24, 40, generator=g).cuda()
    with torch.no_grad():
        got = layer(x)
        ref = torch.relu(layer.norm_layer(layer.conv2d(x)))
    assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max())
