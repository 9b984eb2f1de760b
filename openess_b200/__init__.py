"""openess_b200 -- B200-native (sm_100a CUDA) implementation of the OpenESS per-step hot path.

Layout mirrors the reference tree for the callables that form its drop-in boundary:

    openess_b200.datasets.data_util              <- datasets/data_util.py
    openess_b200.DSEC.dataset.representations    <- DSEC/dataset/representations.py
    openess_b200.utils.loss_functions            <- utils/loss_functions.py
    openess_b200.evaluation.metrics              <- evaluation/metrics.py
    openess_b200.patch.patch_reference()         rebinds the reference's names to these

All compute goes through the C ABI in include/openess_b200.h (openess_b200/lib/libopeness_b200.so,
built from openess_b200/csrc/*.cu).  There is no CPU fallback.
"""
from ._lib import MODE_ATOMIC, MODE_ORDERED, OpenESSB200Error, lib  # noqa: F401

__version__ = "0.1.0"
