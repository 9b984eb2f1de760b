"""Host-to-device prefetch of the NEXT batch on a side stream while the current step computes.

The reference's loop moves a batch with blocking `.to(device)` calls at the top of `train_step`
(training/pretrain_trainer.py:333-344); with raw event slabs (`RawEvents`, ~9 bytes per event instead of a dense tensor) the
copy is ~100 MB per step at B = 4 -- 2 ms at PCIe speed that a copy stream hides completely behind the previous step."""
import torch

from .pretrain_step import RawEvents


class DevicePrefetcher:
    """`feed(batch)` starts the copies of a (pinned) host batch on the copy stream; `take()` returns the device batch fed last,
    ordered after the copies on the current stream.  Batch items may be tensors, `RawEvents`, None, or nested tuples / lists."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._next = None

    def _move(self, item):
        if isinstance(item, torch.Tensor):
            return item.to(self.device, non_blocking=True)
        if isinstance(item, RawEvents):
            return item._replace(**{k: self._move(getattr(item, k)) for k in ("x", "y", "t", "p", "frame_offsets", "rectify_map")})
        if isinstance(item, (tuple, list)):
            return type(item)(self._move(i) for i in item)
        return item

    def _record(self, item, stream):
        if isinstance(item, torch.Tensor):
            if item.is_cuda:
                item.record_stream(stream)
        elif isinstance(item, RawEvents):
            for k in ("x", "y", "t", "p", "frame_offsets", "rectify_map"):
                self._record(getattr(item, k), stream)
        elif isinstance(item, (tuple, list)):
            for i in item:
                self._record(i, stream)

    def feed(self, batch):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))   # not before the consumer of the buffers being recycled
        with torch.cuda.stream(self.stream):
            self._next = self._move(batch)

    def take(self):
        cur, self._next = self._next, None
        if cur is None:
            raise RuntimeError("DevicePrefetcher.take() without a batch fed")
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.stream)
        self._record(cur, main)
        return cur
