"""The C oracle (oracle/oracle.c) vs. golden vectors produced by the reference's own Python code
(oracle/make_golden.py).  CPU only.  Bit-exact for voxelisers / histogram / confusion."""
import hashlib

import numpy as np
import pytest

from conftest import bits_equal, golden_cases, load_golden


def test_tbilinear_bit_exact(oracle):
    z = load_golden("tbilinear")
    names = golden_cases(z)
    assert len(names) >= 12
    for name in names:
        H, W, C = (int(v) for v in z[f"{name}__meta"])
        for sp in (0, 1):
            ev = z[f"{name}__in"].copy()
            out = oracle.voxel_tbilinear(ev, (H, W), C, bool(sp))
            assert bits_equal(out, z[f"{name}__out_sp{sp}"]), (name, sp)
            assert np.array_equal(ev[:, 3], z[f"{name}__pmut"]), f"{name}: polarity mutation side effect"


def test_tbilinear_ddd17_digest(oracle):
    z = load_golden("tbilinear_ddd17")
    H, W, C = (int(v) for v in z["meta"])
    out = oracle.voxel_tbilinear(z["ev"].copy(), (H, W), C, False)
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(z["sha256"])
    assert bits_equal(out.ravel()[z["sample_idx"]], z["sample_val"])


def test_histogram_bit_exact(oracle):
    z = load_golden("histogram")
    for name in golden_cases(z):
        H, W = (int(v) for v in z[f"{name}__meta"])
        ev = z[f"{name}__in"].copy()
        out = oracle.histogram(ev, (H, W))
        assert bits_equal(out, z[f"{name}__out"]), name
        assert np.array_equal(ev[:, 3], z[f"{name}__pmut"])


def test_trilinear_bit_exact(oracle):
    z = load_golden("trilinear")
    names = golden_cases(z)
    assert len(names) >= 14
    for name in names:
        C, H, W, norm = (int(v) for v in z[f"{name}__meta"])
        out = oracle.voxel_trilinear(z[f"{name}__x"], z[f"{name}__y"], z[f"{name}__pol"], z[f"{name}__t"],
                                     C, H, W, normalize=bool(norm))
        ref = z[f"{name}__out"]
        if norm:  # float reductions: tolerance, and the zero pattern must match exactly
            assert np.array_equal(out == 0, ref == 0)
            np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)
        else:
            assert bits_equal(out, ref), name


def test_trilinear_dsec_digest(oracle):
    z = load_golden("trilinear_dsec")
    C, H, W, _ = (int(v) for v in z["meta"])
    out = oracle.voxel_trilinear(z["x"], z["y"], z["pol"], z["t"], C, H, W)
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(z["sha256"])
    assert bits_equal(out.ravel()[z["sample_idx"]], z["sample_val"])


def test_dsec_prestep_bit_exact(oracle):
    z = load_golden("dsec_prestep")
    xo, yo, po, to = oracle.dsec_rectify_tnorm(z["x"], z["y"], z["t"], z["p"], z["rectify_map"])
    for got, want in ((xo, z["xo"]), (yo, z["yo"]), (po, z["po"]), (to, z["to"])):
        assert bits_equal(got, want)


def test_nonzero_standardize(oracle):
    z = load_golden("normalize")
    out, stats = oracle.nonzero_standardize(z["x"])
    assert np.array_equal(out == 0, z["out"] == 0)
    np.testing.assert_allclose(out, z["out"], rtol=2e-5, atol=2e-6)
    outz, _ = oracle.nonzero_standardize(z["zeros"])
    assert bits_equal(outz, z["zeros_out"])


def test_confusion_and_miou(oracle):
    z = load_golden("losses")
    conf = oracle.confusion(z["met__pred"], z["met__gt"], 11, 255)
    assert np.array_equal(conf, z["met__conf"])
    miou, iou, acc = oracle.miou_acc(conf)
    assert miou == pytest.approx(float(z["met__miou"]), rel=1e-12)
    assert acc == pytest.approx(float(z["met__acc"]), rel=1e-12)
    np.testing.assert_allclose(iou, z["met__iou"], rtol=1e-12)


def test_segpool_infonce(oracle):
    z = load_golden("losses")
    S = int(z["pool__S"])
    k, cnt = oracle.segpool(z["pool__feat_voxel"], z["pool__superpixels"], S)
    q, _ = oracle.segpool(z["pool__feat_frame"], z["pool__superpixels"], S)
    assert k.shape == z["pool__k"].shape
    np.testing.assert_allclose(k, z["pool__k"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(q, z["pool__q"], rtol=1e-5, atol=1e-6)
    loss, dk, dq = oracle.infonce(z["pool__k"], z["pool__q"], 0.07, grad=True)
    assert loss == pytest.approx(float(z["pool__nce"]), rel=2e-6)


def test_dice_ce(oracle):
    z = load_golden("losses")
    r = oracle.dice_ce(z["task__logits"], z["task__target"], 255, grad=True)
    assert r["total"] == pytest.approx(float(z["task__total"]), rel=2e-6)
    assert r["dice"] == pytest.approx(float(z["task__dice"]), rel=2e-6)
    np.testing.assert_allclose(r["dlogits"], z["task__dlogits"], rtol=2e-4, atol=1e-9)
