"""ctypes front-end of the CPU parity oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's
`cpu_baseline` / `--impl reference` legs.  Nothing under openess_b200/ may import this module.

Parity pin: the reference has no tests for this path; the oracle is pinned bit-for-bit (voxelisers,
histogram, confusion) / to tolerance (float reductions) against vectors produced by the reference's
own Python code in the build container (oracle/make_golden.py -> tests/golden/*.npz).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty)


def _check(rc, what):
    if rc == -2:
        raise IndexError(f"{what}: empty event array (reference raises IndexError)")
    if rc == -4:
        raise IndexError(f"{what}: index out of range")
    if rc != 0:
        raise ValueError(f"{what}: oracle error {rc}")


def voxel_tbilinear(events, shape, nr_temporal_bins, separate_pol=True):
    """data_util.py:51-117.  `events` [N,4] int64 or float64, C-contiguous; p column mutated in place."""
    H, W = shape
    C = int(nr_temporal_bins)
    assert events.ndim == 2 and events.shape[1] == 4 and events.flags.c_contiguous
    out = np.empty(((2 * C) if separate_pol else C, H, W), np.float32)
    if events.dtype == np.int64:
        rc = lib().oracle_voxel_tbilinear_i64(_p(events, _i64p), ctypes.c_int64(events.shape[0]), C, H, W,
                                              int(separate_pol), _p(out, _f32p))
    elif events.dtype == np.float64:
        rc = lib().oracle_voxel_tbilinear_f64(_p(events, _f64p), ctypes.c_int64(events.shape[0]), C, H, W,
                                              int(separate_pol), _p(out, _f32p))
    else:
        raise TypeError(events.dtype)
    _check(rc, "voxel_tbilinear")
    return out


def histogram(events, shape):
    """data_util.py:17-35."""
    H, W = shape
    assert events.ndim == 2 and events.shape[1] == 4 and events.flags.c_contiguous
    out = np.empty((2, H, W), np.float32)
    if events.dtype == np.int64:
        rc = lib().oracle_histogram_i64(_p(events, _i64p), ctypes.c_int64(events.shape[0]), H, W, _p(out, _f32p))
    elif events.dtype == np.float64:
        rc = lib().oracle_histogram_f64(_p(events, _f64p), ctypes.c_int64(events.shape[0]), H, W, _p(out, _f32p))
    else:
        raise TypeError(events.dtype)
    _check(rc, "histogram")
    return out


def voxel_trilinear(x, y, pol, t, C, H, W, normalize=False, out=None):
    """representations.py:15-55 (serial put_ order).  `out` may be a reusable [C,H,W] float32 buffer."""
    x, y, pol, t = (np.ascontiguousarray(a, np.float32) for a in (x, y, pol, t))
    assert x.shape == y.shape == pol.shape == t.shape and x.ndim == 1
    if out is None:
        out = np.empty((C, H, W), np.float32)
    assert out.shape == (C, H, W) and out.dtype == np.float32 and out.flags.c_contiguous
    rc = lib().oracle_voxel_trilinear(_p(x, _f32p), _p(y, _f32p), _p(pol, _f32p), _p(t, _f32p),
                                      ctypes.c_int64(x.shape[0]), C, H, W, int(bool(normalize)), _p(out, _f32p))
    _check(rc, "voxel_trilinear")
    return out


def dsec_rectify_tnorm(x, y, t, p, rectify_map):
    """sequence_ov.py:204-210 + :154-159 -> (x', y', pol, t_norm) float32."""
    x = np.ascontiguousarray(x, np.uint16)
    y = np.ascontiguousarray(y, np.uint16)
    t = np.ascontiguousarray(t, np.int64)
    p = np.ascontiguousarray(p, np.uint8)
    m = np.ascontiguousarray(rectify_map, np.float32)
    H, W = m.shape[:2]
    n = x.shape[0]
    xo, yo, po, to = (np.empty(n, np.float32) for _ in range(4))
    rc = lib().oracle_dsec_rectify_tnorm(_p(x, _u16p), _p(y, _u16p), _p(t, _i64p), _p(p, _u8p), _p(m, _f32p),
                                         ctypes.c_int64(n), H, W, _p(xo, _f32p), _p(yo, _f32p), _p(po, _f32p),
                                         _p(to, _f32p))
    _check(rc, "dsec_rectify_tnorm")
    return xo, yo, po, to


def nonzero_standardize(x):
    """data_util.py:38-48 / inference_utils.py:77-85.  Returns (normalised copy, [sum, sumsq, nnz])."""
    out = np.ascontiguousarray(x, np.float32).copy()
    stats = np.zeros(3, np.float64)
    rc = lib().oracle_nonzero_standardize(_p(out, _f32p), ctypes.c_int64(out.size), _p(stats, _f64p))
    _check(rc, "nonzero_standardize")
    return out, stats


def confusion(pred, gt, K, ignore):
    """metrics.py:4-23 -> [K,K] int64, conf[gt, pred]."""
    pred = np.ascontiguousarray(pred, np.int64).ravel()
    gt = np.ascontiguousarray(gt, np.int64).ravel()
    conf = np.zeros((K, K), np.int64)
    rc = lib().oracle_confusion(_p(pred, _i64p), _p(gt, _i64p), ctypes.c_int64(pred.size), K,
                                ctypes.c_int64(ignore), _p(conf, _i64p))
    _check(rc, "confusion")
    return conf


def miou_acc(conf):
    """metrics.py:26-36 in float64."""
    conf = conf.astype(np.float64)
    diag = np.diag(conf)
    iou = 100 * diag / np.clip(conf.sum(1) + conf.sum(0) - diag, 1e-12, None)
    acc = 100 * diag.sum() / np.clip(conf.sum(), 1e-12, None)
    return iou.mean(), iou, acc


def segpool(feat, seg, S, M=None):
    """pretrain_trainer.py:445-465 -> (pooled [M,Cf] f32, counts [M] f32)."""
    feat = np.ascontiguousarray(feat, np.float32)
    seg = np.ascontiguousarray(seg, np.int64)
    B, Cf, H, W = feat.shape
    if M is None:
        M = int((seg + np.arange(B)[:, None, None] * S).max()) + 1
    pooled = np.empty((M, Cf), np.float32)
    counts = np.empty((M,), np.float32)
    rc = lib().oracle_segpool(_p(feat, _f32p), _p(seg, _i64p), B, Cf, H, W, int(S), ctypes.c_int64(M),
                              _p(pooled, _f32p), _p(counts, _f32p))
    _check(rc, "segpool")
    return pooled, counts


def infonce(k, q, temperature, grad=False):
    """loss_functions.py:147-153 in float64; optionally (loss, dk, dq)."""
    k = np.ascontiguousarray(k, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    M, D = k.shape
    loss = ctypes.c_double()
    dk = np.zeros((M, D), np.float64) if grad else None
    dq = np.zeros((M, D), np.float64) if grad else None
    rc = lib().oracle_infonce(_p(k, _f32p), _p(q, _f32p), ctypes.c_int64(M), D, ctypes.c_double(temperature),
                              ctypes.byref(loss), _p(dk, _f64p) if grad else None, _p(dq, _f64p) if grad else None)
    _check(rc, "infonce")
    return (loss.value, dk, dq) if grad else loss.value


def dice_ce(logits, target, ignore=255, w_dice=1.0, w_ce=1.0, grad=False):
    """loss_functions.py:17-24,114-135 in float64 -> dict(dice, ce, total[, dlogits])."""
    logits = np.ascontiguousarray(logits, np.float32)
    target = np.ascontiguousarray(target, np.int64)
    B, K, H, W = logits.shape
    out = np.zeros(3, np.float64)
    dl = np.zeros(logits.shape, np.float64) if grad else None
    rc = lib().oracle_dice_ce(_p(logits, _f32p), _p(target, _i64p), B, K, H, W, ctypes.c_int64(ignore),
                              ctypes.c_double(w_dice), ctypes.c_double(w_ce), _p(out, _f64p),
                              _p(dl, _f64p) if grad else None)
    _check(rc, "dice_ce")
    r = {"dice": out[0], "ce": out[1], "total": out[2]}
    if grad:
        r["dlogits"] = dl
    return r
