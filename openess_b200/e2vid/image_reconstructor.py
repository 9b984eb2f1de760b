"""Mirror of e2vid/image_reconstructor.py:19-123 for the OpenESS hot path: same constructor arguments, the same
`last_states_for_each_channel` attribute and `update_reconstruction(event_tensor) -> (img, states, latent)`, but
  * no CudaTimer blocks (the reference synchronises the device 5 times per recurrent step, timers.py:19-26),
  * EventPreprocessor normalisation as two fused kernels without the host-side `num_nonzeros > 0` branch,
  * latent-only encoder by default (img is None; every OpenESS trainer discards it)."""
from math import ceil, floor

import torch
from torch.nn import ReflectionPad2d

from .utils.inference_utils import EventPreprocessor


def optimal_crop_size(max_size, max_subsample_factor):
    """inference_utils.py:270-281."""
    return int(pow(2, max_subsample_factor) * ceil(max_size / pow(2, max_subsample_factor)))


class CropParameters:
    """inference_utils.py:284-311 (padding to a multiple of 2**num_encoders; a no-op at 440x640 and 200x352)."""

    def __init__(self, width, height, num_encoders):
        self.height, self.width, self.num_encoders = height, width, num_encoders
        self.width_crop_size = optimal_crop_size(width, num_encoders)
        self.height_crop_size = optimal_crop_size(height, num_encoders)
        self.padding_top = ceil(0.5 * (self.height_crop_size - height))
        self.padding_bottom = floor(0.5 * (self.height_crop_size - height))
        self.padding_left = ceil(0.5 * (self.width_crop_size - width))
        self.padding_right = floor(0.5 * (self.width_crop_size - width))
        self.is_noop = not (self.padding_top or self.padding_bottom or self.padding_left or self.padding_right)
        self.pad = ReflectionPad2d((self.padding_left, self.padding_right, self.padding_top, self.padding_bottom))
        self.cx, self.cy = floor(self.width_crop_size / 2), floor(self.height_crop_size / 2)
        self.ix0, self.ix1 = self.cx - floor(width / 2), self.cx + ceil(width / 2)
        self.iy0, self.iy1 = self.cy - floor(height / 2), self.cy + ceil(height / 2)


class ImageReconstructor:
    def __init__(self, model, height, width, num_bins, device, options, standardization=False, augmentation=False):
        if standardization or augmentation:
            raise NotImplementedError("image standardisation / augmentation act on the reconstructed image, which the "
                                      "latent-only OpenESS path never produces")
        self.model = model
        self.device = device
        self.height, self.width, self.num_bins = height, width, num_bins
        self.no_recurrent = getattr(options, "no_recurrent", False)
        self.crop = CropParameters(self.width, self.height, self.model.num_encoders)
        self.last_states_for_each_channel = {'grayscale': None}
        self.event_preprocessor = EventPreprocessor(options)

    def update_reconstruction(self, event_tensor, event_tensor_id=None, stamp=None):
        with torch.no_grad():
            events = event_tensor.to(self.device)
            events = self.event_preprocessor(events)
            if not self.crop.is_noop:
                events = self.crop.pad(events)
            out, states, latent = self.model(events, self.last_states_for_each_channel['grayscale'])
            self.last_states_for_each_channel['grayscale'] = None if self.no_recurrent else states
        return out, states, latent
