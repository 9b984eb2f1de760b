"""Drop-in mirror of the reference's datasets/data_util.py (same names, signatures, dtypes, side effects),
computing on the GPU through libopeness_b200.  numpy in -> numpy out, like the reference.

Bit-exact with the reference in the default 'ordered' mode (set OPENESS_B200_VOXEL_MODE=atomic or pass
mode='atomic' for the float-atomics fast path).
"""
import numpy as np
import torch

from .. import voxel as _voxel


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("openess_b200.datasets.data_util needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def generate_input_representation(events, event_representation, shape, nr_temporal_bins=5, separate_pol=True):
    """data_util.py:6-14.  Returns None for unknown representations, like the reference."""
    if event_representation == 'histogram':
        return generate_event_histogram(events, shape)
    elif event_representation == 'voxel_grid':
        return generate_voxel_grid(events, shape, nr_temporal_bins, separate_pol)


def _upload_events(events):
    # The reference accepts any [N, 4] ndarray; its call sites pass int64 (DDD17 memmap,
    # example_loader_ddd17.py:48-52) or float64 (np.stack of mixed columns, sequence_ov.py:268).
    if events.dtype not in (np.int64, np.float64):
        raise TypeError(f"events dtype {events.dtype} not supported (int64 or float64, as the reference's loaders)")
    return torch.from_numpy(np.ascontiguousarray(events)).to(_device(), non_blocking=False)


def generate_event_histogram(events, shape):
    """data_util.py:17-35 -> float32 [2, H, W] = stack([neg, pos]); mutates events[:, 3] (0 -> -1)."""
    height, width = shape
    p = events[:, 3]
    p[p == 0] = -1                                    # side effect on the caller's array, data_util.py:26
    ev = _upload_events(events)
    status = torch.zeros(1, dtype=torch.int32, device=ev.device)
    out = _voxel.voxel_histogram(ev, height, width, mutate_p=False, status=status)
    if int(status.item()):
        raise IndexError("event index out of bounds for the histogram image")   # numpy raises IndexError too
    return out[0].cpu().numpy()


def normalize_voxel_grid(events):
    """data_util.py:38-48 on a torch tensor (CPU or CUDA): nonzero mean / biased std standardisation."""
    x = events.detach().to(_device(), dtype=torch.float32, copy=True).contiguous()
    _voxel.nonzero_standardize(x, n_groups=1, unbiased=False)
    return x.to(events.device)


def generate_voxel_grid(events, shape, nr_temporal_bins, separate_pol=True, mode=None):
    """data_util.py:51-117 -> float32 [C or 2C, H, W]; mutates events[:, 3] (0 -> -1)."""
    height, width = shape
    assert(events.shape[1] == 4)
    assert(nr_temporal_bins > 0)
    assert(width > 0)
    assert(height > 0)
    if events.shape[0] == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")    # events[-1, 2], data_util.py:67
    pols = events[:, 3]
    pols[pols == 0] = -1                              # side effect on the caller's array, data_util.py:78-79
    ev = _upload_events(events)
    out = _voxel.voxel_tbilinear(ev, int(nr_temporal_bins), height, width, separate_pol=separate_pol, mode=mode,
                                 mutate_p=False)
    return out[0].cpu().numpy()
