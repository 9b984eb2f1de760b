import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The goldens are the reference's fp32 CPU outputs: keep the cuDNN / cuBLAS layers that still run under the
    # mirrors in true fp32 (torch's default lets cuDNN convolutions use TF32, ~1e-3 relative error).
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_cases(z, keys_suffix="__meta"):
    return sorted({k.split("__")[0] for k in z.files if k.endswith(keys_suffix)})


def bits_equal(a, b):
    """Bit-exact comparison that also treats NaNs with identical payloads as equal."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
