"""ctypes binding of libopeness_b200.so (include/openess_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopeness_b200.so")

MODE_ORDERED = 0
MODE_ATOMIC = 1
KIND_TRILINEAR = 0
KIND_TBILINEAR = 1

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t
_f32 = ctypes.c_float

# name -> argtypes (restype is always int unless noted); mirrors include/openess_b200.h one to one
SIGNATURES = {
    "oess_abi_version": [],
    "oess_error_string": [_int],
    "oess_launch_count": [],
    "oess_profile_begin": [],
    "oess_profile_end": [ctypes.c_char_p, _sz],
    "oess_voxel_ws_bytes": [_int, _int, _i64, _int, _int, _int, _int, ctypes.POINTER(_sz)],
    "oess_voxel_trilinear": [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp],
    "oess_voxel_tbilinear_i64": [_vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp],
    "oess_voxel_tbilinear_f64": [_vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp],
    "oess_voxel_histogram_i64": [_vp, _vp, _i64, _int, _int, _int, _int, _vp, _vp, _vp],
    "oess_voxel_histogram_f64": [_vp, _vp, _i64, _int, _int, _int, _int, _vp, _vp, _vp],
    "oess_voxel_tbilinear_ddd17": [_vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp],
    "oess_voxel_histogram_ddd17": [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp],
    "oess_dsec_rectify_tnorm": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "oess_dsec_rectify_tnorm_u32": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "oess_nonzero_standardize": [_vp, _i64, _int, _vp, _int, _int, _vp],
    "oess_segpool_ws_bytes": [_int, _int, _int, _int, _i64, ctypes.POINTER(_sz)],
    "oess_segpool_fwd": [_vp, _vp, _int, _int, _int, _int, _int, _i64, _vp, _vp, _vp, _vp, _sz, _vp],
    "oess_segpool_bwd": [_vp, _vp, _vp, _int, _int, _int, _int, _int, _i64, _vp, _vp],
    "oess_infonce_ws_bytes": [_i64, _int, ctypes.POINTER(_sz)],
    "oess_infonce": [_vp, _vp, _i64, _int, _f32, _vp, _vp, _vp, _vp, _sz, _vp],
    "oess_dice_ce_partials": [_vp, _vp, _int, _int, _int, _int, _i64, _vp, _vp],
    "oess_dice_ce_finish": [_vp, _int, _f32, _f32, _vp, _vp],
    "oess_dice_ce_finish_ex": [_vp, _int, _i64, _f32, _f32, _vp, _vp],
    "oess_dice_ce_bwd": [_vp, _vp, _int, _int, _int, _int, _i64, _vp, _f32, _f32, _vp, _vp, _vp],
    "oess_confusion": [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp],
    "oess_argmax_confusion": [_vp, _vp, _int, _int, _int, _int, _i64, _vp, _vp, _vp],
    "oess_convlstm_gates": [_vp, _vp, _vp, _vp, _int, _int, _i64, _vp],
    "oess_l1_mean": [_vp, _vp, _i64, _vp, _vp, _vp],
    "oess_l1_mean_bwd": [_vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "oess_cos_consistency": [_vp, _vp, _int, _int, _i64, _vp, _vp, _vp],
    "oess_cos_consistency_bwd": [_vp, _vp, _int, _int, _i64, _vp, _vp, _vp, _vp],
    "oess_pixel_linear": [_vp, _vp, _vp, _int, _int, _int, _i64, _vp, _vp],
    "oess_pixel_linear_wgrad_ws_bytes": [_int, _int, ctypes.POINTER(_sz)],
    "oess_pixel_linear_wgrad": [_vp, _vp, _int, _int, _int, _i64, _vp, _vp, _vp, _sz, _vp],
    "oess_gemm_tf32": [_vp, _vp, _vp, _vp, _i64, _int, _int, _vp],
    "oess_gemm_tf32_ex": [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp],
    "oess_hflip_rows": [_vp, _int, _int, _i64, _int, _vp, _vp],
    "oess_frame_color_aug": [_vp, _int, _i64, _vp, _vp, _vp, _vp, _vp],
    "oess_unsharp_rescale": [_vp, _vp, _int, _int, _int, _f32, _f32, _f32, _int, _vp, _vp],
    "oess_zero_insert2x_nhwc": [_vp, _vp, _int, _int, _int, _int, _vp, _vp],
    "oess_bilinear_resize_planes": [_vp, _i64, _int, _int, _int, _int, _vp, _vp],
    "oess_bilinear_resize_planes_bwd": [_vp, _i64, _int, _int, _int, _int, _vp, _vp, _vp],
    "oess_upsample2x_cat_nhwc": [_vp, _vp, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_upsample2x_cat_nhwc_bwd": [_vp, _int, _int, _int, _int, _int, _vp, _vp, _vp],
    "oess_pred_sigmoid_nhwc": [_vp, _vp, _vp, _f32, _i64, _int, _vp, _vp],
    "oess_maxpool3x3s2_nhwc": [_vp, _int, _int, _int, _int, _vp, _vp],
    "oess_global_avgpool_nhwc": [_vp, _int, _i64, _int, _vp, _vp],
    "oess_vit_patchify": [_vp, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_vit_assemble": [_vp, _vp, _vp, _int, _int, _int, _vp, _vp],
    "oess_layernorm_rows": [_vp, _vp, _vp, _f32, _i64, _int, _vp, _vp],
    "oess_mha_fwd": [_vp, _int, _int, _int, _vp, _vp],
    "oess_mha_fwd_tc": [_vp, _int, _int, _int, _vp, _vp],
    "oess_l2norm_rows": [_vp, _i64, _int, _vp],
    "oess_bilinear_tokens_to_nchw": [_vp, _int, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_convlstm_step_nhwc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp],
    "oess_convlstm_step_nhwc_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp],
    "oess_conv2d_nhwc_tf32_bf16out": [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp],
    "oess_upnorm_pool_fwd": [_vp, _vp, _int, _int, _int, _int, _int, _int, _int, _i64, _vp, _vp, _vp, _vp],
    "oess_upnorm_pool_bwd": [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _i64, _vp, _vp],
    "oess_conv2d_nhwc_tf32_stats": [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_batchnorm_nhwc_sums": [_vp, _i64, _int, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _int, _vp, _sz, _vp],
    "oess_conv2d_nhwc_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_batchnorm_nhwc_sums_bf16": [_vp, _i64, _int, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _int, _vp, _int, _vp, _sz, _vp],
    "oess_planes_to_nhwc_padded": [_vp, _int, _int, _i64, _vp, _int, _vp, _vp],
    "oess_planes_to_nhwc_padded_w": [_vp, _int, _int, _int, _int, _vp, _int, _int, _vp, _vp],
    "oess_conv2d_nhwc_tf32_rowunfold": [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _vp],
    "oess_conv2d_nhwc_tf32_instats": [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _vp],
    "oess_instancenorm_nhwc_sums": [_vp, _int, _i64, _int, _vp, _f32, _vp, _int, _vp],
    "oess_conv2d_wgrad_nhwc_tf32": [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp],
    "oess_instancenorm_nhwc_sums_train": [_vp, _int, _i64, _int, _vp, _f32, _vp, _int, _vp, _vp],
    "oess_instancenorm_nhwc_bwd": [_vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _f32, _vp, _vp, _vp],
    "oess_batchnorm_nhwc_sums_train": [_vp, _i64, _int, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _int, _vp, _vp, _sz, _vp],
    "oess_batchnorm_nhwc_bwd": [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _f32, _vp, _vp, _vp],
    "oess_bn_ws_bytes": [_int, ctypes.POINTER(_sz)],
    "oess_batchnorm_nhwc": [_vp, _i64, _int, _vp, _vp, _vp, _vp, _f32, _f32, _int, _vp, _int, _vp, _sz, _vp],
    "oess_conv2d_nhwc_tf32": [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _int, _vp],
}

_lib = None
_lock = threading.Lock()


class OpenESSB200Error(RuntimeError):
    pass


def lib():
    """Load libopeness_b200.so.  Raises if it has not been built (python -m openess_b200.build)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise OpenESSB200Error(
                        f"{LIB_PATH} not found: build it with `python -m openess_b200.build` "
                        "(openess_b200 has no CPU fallback)")
                h = ctypes.CDLL(LIB_PATH)
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(h, name)  # AttributeError if the header and the library disagree
                    fn.argtypes = argtypes
                    fn.restype = {"oess_error_string": ctypes.c_char_p,
                                  "oess_launch_count": ctypes.c_ulonglong}.get(name, ctypes.c_int)
                _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().oess_error_string(rc).decode()
        raise OpenESSB200Error(f"{what} failed: {msg} (code {rc})")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OpenESSB200Error("openess_b200 kernels need CUDA tensors (no CPU fallback)")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream).  Kernels using it are ordered on that stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def voxel_ws_bytes(kind, mode, n, F, C, H, W):
    out = _sz(0)
    check(lib().oess_voxel_ws_bytes(kind, mode, n, F, C, H, W, ctypes.byref(out)), "oess_voxel_ws_bytes")
    return out.value


def launch_count():
    """Kernels launched by libopeness_b200 so far in this process."""
    return int(lib().oess_launch_count())


class profile:
    """Context manager: per-kernel CUDA-event timing of every launch made by this host thread.

    with profile() as p: ...; p.kernels -> {name: (launches, total_ms)}"""

    def __enter__(self):
        check(lib().oess_profile_begin(), "oess_profile_begin")
        self.kernels = {}
        return self

    def __exit__(self, *exc):
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().oess_profile_end(buf, len(buf)), "oess_profile_end")
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(",", 2)
            self.kernels[name] = (int(cnt), float(ms))
        return False
