"""GPU augmentation against the reference's own operations: torch.flip and
torchvision.transforms.functional.adjust_brightness / adjust_contrast (what sequence_ov.py:387-407 calls), and the replay of
the reference's `random` call sequence.  Flips are exact; the colour chain agrees to 2e-6 (float64 mean of the grey image
instead of torch.mean's float32 cascade)."""
import random

import pytest
import torch


def test_draw_params_replays_the_reference_random_sequence():
    from openess_b200.DSEC.dataset.augment import draw_params
    r1, r2 = random.Random(1205), random.Random(1205)
    got = draw_params(5, r1)
    for b in range(5):                                            # sequence_ov.py:388-406, literally
        assert got["flip"][b] == (r2.random() >= 0.5)
        bf = r2.uniform(0.8, 1.2) if r2.random() >= 0.5 else 1.0
        cf = r2.uniform(0.8, 1.2) if r2.random() >= 0.5 else 1.0
        assert got["brightness"][b] == bf and got["contrast"][b] == cf
        assert got["noise"][b] == (r2.random() >= 0.5)
    assert any(got["flip"]) and not all(got["flip"])
    from openess_b200.DSEC.dataset import augment
    with pytest.raises(Exception):
        augment.hflip_rows_(torch.zeros(2, 3, 4), torch.zeros(2, dtype=torch.uint8))     # CPU tensors: no fallback


@pytest.mark.gpu
def test_augment_batch_matches_torch_and_torchvision():
    import torchvision.transforms.functional as TF
    from openess_b200.DSEC.dataset.augment import augment_batch_, frame_color_aug_, hflip_rows_
    g = torch.Generator(device="cuda").manual_seed(4)
    B, H, W = 4, 44, 70                                           # odd half-width exercises the middle column
    event = torch.randn(B, 10, H, W, device="cuda", generator=g)
    frame = torch.rand(B, 3, H, W, device="cuda", generator=g)
    label = torch.randint(0, 11, (B, H, W), device="cuda", generator=g)
    pl = torch.randint(0, 256, (B, H, W), device="cuda", generator=g)
    sp = torch.randint(0, 100, (B, H, W), device="cuda", generator=g)
    params = {"flip": [True, False, True, False], "brightness": [1.13, 1.0, 0.85, 1.2], "contrast": [0.9, 1.17, 1.0, 1.2],
              "noise": [False, False, False, False]}
    want_e, want_f, want_l, want_p, want_s = [], [], [], [], []
    for b in range(B):
        e, f, l, p_, s = event[b].cpu(), frame[b].cpu(), label[b].cpu(), pl[b].cpu(), sp[b].cpu()
        if params["flip"][b]:
            e, l, f, p_, s = torch.flip(e, [2]), torch.flip(l, [1]), torch.flip(f, [2]), torch.flip(p_, [1]), torch.flip(s, [1])
        if params["brightness"][b] != 1.0:
            f = TF.adjust_brightness(f, params["brightness"][b])
        if params["contrast"][b] != 1.0:
            f = TF.adjust_contrast(f, params["contrast"][b])
        want_e.append(e); want_f.append(f); want_l.append(l); want_p.append(p_); want_s.append(s)
    e2, l2, f2, p2, s2 = augment_batch_(event.clone(), label.clone(), frame.clone(), pl.clone(), sp.clone(), params)
    assert torch.equal(e2.cpu(), torch.stack(want_e)) and torch.equal(l2.cpu(), torch.stack(want_l))
    assert torch.equal(p2.cpu(), torch.stack(want_p)) and torch.equal(s2.cpu(), torch.stack(want_s))
    err = float((f2.cpu() - torch.stack(want_f)).abs().max())
    print("colour chain max |err| vs torchvision: %.2e" % err)
    assert err < 2e-6
    # noise: added after the clamps, only where gated
    noise = torch.randn(frame.shape, device="cuda", generator=g) * 0.05
    f3 = frame_color_aug_(frame.clone(), torch.ones(B), torch.ones(B), noise)
    assert torch.equal(f3, frame + noise)
    params["noise"] = [True, False, False, True]
    f4 = augment_batch_(None, None, frame.clone(), None, None, {**params, "flip": [False] * B, "brightness": [1.0] * B,
                                                                  "contrast": [1.0] * B}, generator=g)[2]
    d = (f4 - frame).flatten(1).abs().amax(1).cpu()
    assert d[0] > 0 and d[3] > 0 and d[1] == 0 and d[2] == 0
    # odd width and int32 / float64 element sizes
    x = torch.arange(2 * 3 * 7, device="cuda", dtype=torch.float64).view(2, 3, 7)
    assert torch.equal(hflip_rows_(x.clone(), torch.tensor([1, 0], dtype=torch.uint8, device="cuda")),
                       torch.stack((torch.flip(x[0], [1]), x[1])))
