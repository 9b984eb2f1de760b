"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol include/openess_b200.h
declares, host-side queries work without a GPU, and the product path fails loudly without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from openess_b200 import build, _lib
    build.build_library()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "openess_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oess_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/openess_b200.h but not exported"


def test_binding_covers_header(lib):
    from openess_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_errors(lib):
    assert lib.oess_abi_version() == 1
    assert lib.oess_error_string(0) == b"ok"
    assert b"workspace" in lib.oess_error_string(-2)


def test_ws_query_is_host_only(lib):
    from openess_b200 import _lib
    big = _lib.voxel_ws_bytes(_lib.KIND_TRILINEAR, _lib.MODE_ORDERED, 100000, 1, 5, 480, 640)
    # ATOMIC on sparse frames runs the sort-based pipeline (density dispatch): same workspace; dense frames: atomics kernel
    assert _lib.voxel_ws_bytes(_lib.KIND_TRILINEAR, _lib.MODE_ATOMIC, 100000, 1, 5, 480, 640) == big
    small = _lib.voxel_ws_bytes(_lib.KIND_TRILINEAR, _lib.MODE_ATOMIC, 1000000, 1, 5, 480, 640)
    assert 0 < small < _lib.voxel_ws_bytes(_lib.KIND_TRILINEAR, _lib.MODE_ORDERED, 1000000, 1, 5, 480, 640)
    # ordered: two float4 record buffers + radix histograms (+ per-cell CSR on the generic tall-sensor path)
    assert big >= 2 * 16 * 100000
    tall = _lib.voxel_ws_bytes(_lib.KIND_TRILINEAR, _lib.MODE_ORDERED, 100000, 1, 5, 2000, 640)
    assert tall >= 2 * 16 * 100000 + 4 * 2001 * 641
    tb = _lib.voxel_ws_bytes(_lib.KIND_TBILINEAR, _lib.MODE_ORDERED, 50000, 1, 5, 260, 346)
    assert tb >= 2 * 8 * 50000
    out = ctypes.c_size_t(0)
    assert lib.oess_voxel_ws_bytes(7, 0, 10, 1, 5, 4, 4, ctypes.byref(out)) == -1      # bad kind
    assert lib.oess_voxel_ws_bytes(0, 0, 10, 1, 0, 4, 4, ctypes.byref(out)) == -1      # C <= 0
    assert lib.oess_voxel_ws_bytes(0, 3, 10, 1, 5, 4, 4, ctypes.byref(out)) == -1      # bad mode


def test_argument_errors_without_gpu(lib):
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.oess_voxel_trilinear(None, None, None, None, None, 10, 1, 5, 4, 4, 0, 0, None, None, 0, None) == -1
    assert lib.oess_infonce(None, None, 8, 256, 0.07, None, None, None, None, 0, None) == -1
    assert lib.oess_confusion(None, None, 10, 0, 255, None, None, None) == -1
    sz = ctypes.c_size_t(0)
    assert lib.oess_infonce_ws_bytes(100, 256, ctypes.byref(sz)) == 0 and sz.value >= 800


def test_argument_errors_of_later_rows_without_gpu(lib):
    """Entry points added for rows a14 / a14' / 8f: shape and pointer checks come before any CUDA call."""
    assert lib.oess_layernorm_rows(None, None, None, 1e-6, 10, 100, None, None) == -1          # D % 128 != 0
    assert lib.oess_layernorm_rows(None, None, None, 1e-6, 10, 2048, None, None) == -1         # D > 1024
    assert lib.oess_layernorm_rows(None, None, None, 1e-6, 0, 768, None, None) == 0            # empty: nothing to do
    assert lib.oess_l2norm_rows(None, 5, 512, None) == -1                                      # null pointer
    assert lib.oess_mha_fwd(None, 1, 0, 12, None, None) == -1 and lib.oess_mha_fwd_tc(None, 1, 0, 12, None, None) == -1
    assert lib.oess_mha_fwd_tc(None, 0, 7, 12, None, None) == 0                                # B == 0
    assert lib.oess_gemm_tf32_ex(None, None, None, None, None, 4, 4, 6, 0, None) == -1          # K % 4 != 0 / null
    assert lib.oess_gemm_tf32_ex(None, None, None, None, None, 4, 4, 8, 7, None) == -1          # unknown epilogue flag
    assert lib.oess_vit_patchify(None, 1, 3, 8, 8, 0, None, None) == -1                        # patch <= 0
    assert lib.oess_maxpool3x3s2_nhwc(None, 1, 8, 8, 6, None, None) == -1                      # C % 4 != 0
    assert lib.oess_zero_insert2x_nhwc(None, None, 1, 4, 4, 6, None, None) == -1
    assert lib.oess_pred_sigmoid_nhwc(None, None, None, 0.0, 10, 128, None, None) == -1        # C > 64
    assert lib.oess_hflip_rows(None, 2, 1, 4, 4, None, None) == -1                             # element size 2
    assert lib.oess_frame_color_aug(None, 1, 16, None, None, None, None, None) == -1
    assert lib.oess_voxel_tbilinear_ddd17(None, None, None, 10, 1, 5, 4, 4, 0, 0, None, None, 0, None) == -1
    assert lib.oess_voxel_histogram_ddd17(None, None, None, 0, 0, 4, 4, None, None, None) == 0  # F == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from openess_b200.datasets import data_util
    from openess_b200.DSEC.dataset.representations import VoxelGrid
    ev = np.array([[1, 1, 0, 1], [2, 2, 10, 0]], np.int64)
    with pytest.raises(RuntimeError):
        data_util.generate_voxel_grid(ev, (4, 4), 5, False)
    with pytest.raises(RuntimeError):
        VoxelGrid(5, 4, 4, False).convert(*(torch.zeros(3) for _ in range(4)))
    from openess_b200 import ops
    from openess_b200.DSEC.dataset import augment
    from openess_b200.datasets.extract_data_tools import example_loader_ddd17 as ld
    for call in (lambda: ops.mha_fwd(torch.zeros(4, 384), 1, 4, 2), lambda: ops.layernorm_rows(torch.zeros(2, 128), torch.ones(128), torch.zeros(128), 1e-6),
                 lambda: ops.maxpool3x3s2_nhwc(torch.zeros(1, 4, 4, 4)), lambda: ops.zero_insert2x_nhwc(torch.zeros(1, 4, 2, 2)),
                 lambda: augment.frame_color_aug_(torch.zeros(1, 3, 2, 2), torch.ones(1), torch.ones(1)),
                 lambda: ld.event_tensors(torch.zeros(4, dtype=torch.int64), torch.zeros((4, 3), dtype=torch.int16), None, (4, 4))):
        with pytest.raises(RuntimeError):
            call()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under openess_b200/ may reference it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "openess_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", txt, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
