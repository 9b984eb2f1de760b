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
    def __init__(self, model, height, width, num_bins, device, options, augmentation=False, standardization=False):
        # keyword order as in the reference (image_reconstructor.py:20)
        if augmentation:
            raise NotImplementedError("the albumentations image augmentation (image_reconstructor.py:33-47, 115-121) is a CPU / PIL "
                                      "round trip on the reconstructed image; no OpenESS trainer enables it")
        self.standardization = standardization
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
            if self.standardization:                             # :108-113 per-sample min / max rescale of the image
                if out is None:
                    raise RuntimeError("standardization=True needs the reconstructed image: build E2VIDRecurrent(config, latent_only=False)")
                b, hh, ww = out.size(0), out.size(2), out.size(3)
                out = out.reshape(b, -1)
                out = out - out.min(1, keepdim=True)[0]
                out = out / out.max(1, keepdim=True)[0]
                out = out.view(b, 1, hh, ww)
        return out, states, latent


class PostProcessor:
    """image_reconstructor.py:126-140 + inference_utils.py:90-129, 234-252: what e2vid/run_reconstruction.py applies to every
    reconstruction before it is written as the PNG the `frame2recon` configs read: unsharp mask (5 x 5 Gaussian, sigma
    `unsharp_mask_sigma`, amount `unsharp_mask_amount`), intensity rescaling to [Imin, Imax] (optionally the running median
    of the clipped per-image min / max: `auto_hdr`), clamp, 8-bit quantisation.  One fused kernel per image batch
    (`oess_unsharp_rescale`); the bilateral filter (off by default: `bilateral_filter_sigma` 0) is not built."""

    def __init__(self, device, options):
        from collections import deque
        self.device = device
        self.amount = float(getattr(options, "unsharp_mask_amount", 0.3))
        self.sigma = float(getattr(options, "unsharp_mask_sigma", 1.0))
        self.auto_hdr = bool(getattr(options, "auto_hdr", False))
        self.median_size = int(getattr(options, "auto_hdr_median_filter_size", 10))
        self.Imin, self.Imax = float(getattr(options, "Imin", 0.0)), float(getattr(options, "Imax", 1.0))
        if getattr(options, "bilateral_filter_sigma", 0.0):
            raise NotImplementedError("bilateral filter (cv2, CPU) is not part of the GPU post-processing")
        self.intensity_bounds = deque()
        self.kernel = gaussian_kernel_5x5(self.sigma).to(device)

    def process(self, img):
        import numpy as np
        from .. import ops
        with torch.no_grad():
            img = img.to(self.device).contiguous()
            if self.auto_hdr:
                # the bounds are taken on the SHARPENED image (the rescaler runs after the unsharp mask, :133-134)
                sharp = ops.unsharp_rescale(img, self.kernel, self.amount, 0.0, 1.0, quantize=False)
                lo = float(np.clip(sharp.min().item(), 0.0, 0.45))
                hi = float(np.clip(sharp.max().item(), 0.55, 1.0))
                if len(self.intensity_bounds) > self.median_size:
                    self.intensity_bounds.popleft()
                self.intensity_bounds.append((lo, hi))
                self.Imin = float(np.median([a for a, _ in self.intensity_bounds]))
                self.Imax = float(np.median([b for _, b in self.intensity_bounds]))
            return ops.unsharp_rescale(img, self.kernel, self.amount, self.Imin, self.Imax, quantize=True)


def gaussian_kernel_5x5(sigma):
    """inference_utils.py gkern(5, sigma): outer product of the differences of the normal CDF on a 5-point grid, normalised."""
    import numpy as np
    import scipy.stats as st
    kernlen = 5
    interval = (2 * sigma + 1.) / kernlen
    x = np.linspace(-sigma - interval / 2., sigma + interval / 2., kernlen + 1)
    kern1d = np.diff(st.norm.cdf(x))
    kernel_raw = np.sqrt(np.outer(kern1d, kern1d))
    return torch.from_numpy(kernel_raw / kernel_raw.sum()).float()
