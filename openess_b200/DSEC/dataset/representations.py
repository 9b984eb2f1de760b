"""Drop-in mirror of the reference's DSEC/dataset/representations.py (VoxelGrid), on the GPU."""
import torch

from ... import voxel as _voxel


class EventRepresentation:
    def convert(self, x: torch.Tensor, y: torch.Tensor, pol: torch.Tensor, time: torch.Tensor):
        raise NotImplementedError


class VoxelGrid(EventRepresentation):
    """representations.py:9-55.  Re-entrant: no per-instance scratch state (the reference is called from
    8 joblib threads concurrently, sequence_ov.py:304-305)."""

    def __init__(self, channels: int, height: int, width: int, normalize: bool, mode=None):
        self.nb_channels = channels
        self.height = height
        self.width = width
        self.normalize = normalize
        self.mode = mode

    def convert(self, x: torch.Tensor, y: torch.Tensor, pol: torch.Tensor, time: torch.Tensor):
        assert x.shape == y.shape == pol.shape == time.shape
        assert x.ndim == 1
        if x.numel() == 0:
            raise IndexError("index 0 is out of bounds for dimension 0 with size 0")   # t_norm[0], :25
        src = pol.device                      # the reference returns the grid on pol.device (:21)
        if not torch.cuda.is_available():
            raise RuntimeError("openess_b200 VoxelGrid needs a CUDA device (no CPU fallback)")
        dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        with torch.no_grad():
            args = [a.detach().to(device=dev, dtype=torch.float32).contiguous() for a in (x, y, pol, time)]
            out = _voxel.voxel_trilinear(*args, self.nb_channels, self.height, self.width, mode=self.mode,
                                         normalize=self.normalize)[0]
        return out.to(src)
