"""DDD17 event ingest (SURVEY.md 8f row 1, DDD17 half): mirror of datasets/extract_data_tools/example_loader_ddd17.py:9-54
(`load_files_in_directory`, `load_events`, `extract_events_from_memmap`) plus the B200 wire format behind it.

The reference memory-maps `events.dat.t` (int64 [N, 1]) and `events.dat.xyp` (int16 [N, 3]); per sample it assembles an
int64 [n, 4] array on the host (np.concatenate + widening + column reorder, 32 B / event) which `DDD17Events.__getitem__`
(datasets/ddd17_events_loader.py:146-177) cuts into `nr_events_data` chunks and voxelises one by one with np.add.at.

Here the on-disk records ARE the wire format: `DDD17Stager` copies the record ranges of a batch of samples from the memory
maps into two pinned host buffers (14 B / event, no widening), issues two asynchronous host-to-device copies on the current
stream, and `event_tensors` voxelises every chunk of every sample in ONE launch of `oess_voxel_tbilinear_ddd17` /
`oess_voxel_histogram_ddd17` (bit-identical to the reference: the kernels read the same integers).  The three reference
functions keep their names, signatures and return values for callers that want the host arrays.
There is no CPU voxeliser here: `event_tensors` needs CUDA and the built library.
"""
import glob
import os

import numpy as np
import torch

from ... import voxel as _voxel


def load_files_in_directory(directory, t_interval=50):
    """example_loader_ddd17.py:9-29."""
    name = {10: "index/index_10ms.npy", 50: "index/index_50ms.npy", 250: "index/index_250ms.npy"}.get(t_interval, "index/index_50ms.npy")
    img_timestamp_event_idx = np.load(os.path.join(directory, name))
    t_events, xyp_events = load_events(os.path.join(directory, "events.dat.t"), os.path.join(directory, "events.dat.xyp"))
    segmentation_mask_files = sorted(glob.glob(os.path.join(directory, "segmentation_masks", "*.png")))
    return img_timestamp_event_idx, t_events, xyp_events, segmentation_mask_files


def load_events(t_file, xyp_file):
    """example_loader_ddd17.py:32-38."""
    num_events = int(os.path.getsize(t_file) / 8)
    t_events = np.memmap(t_file, dtype="int64", mode="r", shape=(num_events, 1))
    xyp_events = np.memmap(xyp_file, dtype="int16", mode="r", shape=(num_events, 3))
    return t_events, xyp_events


def event_range(img_idx, img_timestamp_event_idx, fixed_duration=False, nr_events=32000):
    """[begin, end) record range of one sample (example_loader_ddd17.py:44-49)."""
    if fixed_duration:
        _, event_idx, event_idx_before = img_timestamp_event_idx[img_idx]
        event_idx_before = max([event_idx_before, 0])
    else:
        _, event_idx, _ = img_timestamp_event_idx[img_idx]
        event_idx_before = max([event_idx - nr_events, 0])
    return int(event_idx_before), int(event_idx)


def extract_events_from_memmap(t_events, xyp_events, img_idx, img_timestamp_event_idx, fixed_duration=False, nr_events=32000):
    """example_loader_ddd17.py:41-54: host int64 [n, 4] rows (x, y, t, p) -- kept for drop-in callers."""
    b, e = event_range(img_idx, img_timestamp_event_idx, fixed_duration, nr_events)
    ev = np.concatenate([np.array(t_events[b:e], dtype="int64"), np.array(xyp_events[b:e], dtype="int64")], -1)
    return ev[:, [1, 2, 0, 3]]


def chunk_cuts(t_events, begin, end, nr_events_data, fixed_duration):
    """Chunk boundaries of one sample relative to `begin` (datasets/ddd17_events_loader.py:153-170): equal event counts, or
    equal durations located with np.searchsorted on the sample's timestamps."""
    n = end - begin
    if n <= 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")        # t_ns[-1] on an empty sample (:154)
    cuts = [0]
    if fixed_duration:
        t_ns = np.asarray(t_events[begin:end]).reshape(-1)
        delta = int((t_ns[-1] - t_ns[0]) / nr_events_data)
        for i in range(nr_events_data):
            cuts.append(min(int(np.searchsorted(t_ns, t_ns[0] + (i + 1) * delta)), n))
    else:
        per = n // nr_events_data
        for i in range(nr_events_data):
            cuts.append(min(cuts[-1] + per, n))
    return np.asarray(cuts, dtype=np.int64)


class DDD17Stager:
    """Pinned host staging of raw DDD17 records + asynchronous upload.  One instance per loader thread / stream."""

    def __init__(self, capacity_events, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DDD17Stager stages records for the CUDA voxeliser (no CPU path)")
        self._done = None
        self._alloc(int(capacity_events))

    def _alloc(self, cap):
        self.capacity = cap
        self.t_host = torch.empty(cap, dtype=torch.int64).pin_memory()
        self.xyp_host = torch.empty((cap, 3), dtype=torch.int16).pin_memory()
        self.t_dev = torch.empty(cap, dtype=torch.int64, device=self.device)
        self.xyp_dev = torch.empty((cap, 3), dtype=torch.int16, device=self.device)

    def stage(self, t_events, xyp_events, ranges):
        """ranges: [(begin, end), ...] record ranges (one per sample).  Returns (t_dev [n], xyp_dev [n, 3], sample_offsets
        [len(ranges) + 1] int64 numpy): device views valid until the next stage() call on this stager."""
        total = sum(e - b for b, e in ranges)
        if total > self.capacity:
            torch.cuda.current_stream(self.device).synchronize()
            self._alloc(max(total, 2 * self.capacity))
        else:
            if self._done is not None:            # the previous upload must have left the pinned buffers
                self._done.synchronize()
        th, xh = self.t_host.numpy(), self.xyp_host.numpy()
        offs = [0]
        for b, e in ranges:
            o = offs[-1]
            th[o:o + e - b] = np.asarray(t_events[b:e]).reshape(-1)
            xh[o:o + e - b] = xyp_events[b:e]
            offs.append(o + e - b)
        with torch.cuda.device(self.device):
            self.t_dev[:total].copy_(self.t_host[:total], non_blocking=True)
            self.xyp_dev[:total].copy_(self.xyp_host[:total], non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record()
        return self.t_dev[:total], self.xyp_dev[:total], np.asarray(offs, dtype=np.int64)

    @property
    def h2d_bytes_per_event(self):
        return 14


def event_tensors(t_dev, xyp_dev, frame_offsets, shape, event_representation="voxel_grid", nr_temporal_bins=5,
                  separate_pol=True, mode=None):
    """All chunk frames of a staged batch in one launch -> [F, planes, H, W] float32 on the device
    (= generate_input_representation per chunk, datasets/data_util.py:6-14, on the rows the reference would assemble)."""
    H, W = shape
    if event_representation == "histogram":
        return _voxel.voxel_histogram_ddd17(t_dev, xyp_dev, H, W, frame_offsets)
    if event_representation == "voxel_grid":
        return _voxel.voxel_tbilinear_ddd17(t_dev, xyp_dev, nr_temporal_bins, H, W, frame_offsets, separate_pol, mode)
    return None                                                                        # data_util.py:14 falls through


def load_event_tensors(stager, t_events, xyp_events, img_indices, img_timestamp_event_idx, shape, nr_events_data=5,
                       nr_events=160000, fixed_duration=False, event_representation="voxel_grid", nr_temporal_bins=5,
                       separate_pol=True, mode=None):
    """Batch version of the event branch of DDD17Events.__getitem__ (datasets/ddd17_events_loader.py:146-193, before the
    optional resize / crop): -> [B, nr_events_data * planes, H, W], chunk-major like the reference's torch.cat(dim=0).
    `nr_events` is the per-sample total the reference passes to extract_events_from_memmap (:151, = nr_events_data *
    nr_events_per_data).  Records behind the last chunk boundary belong to no chunk in the reference (:161-170: n % chunks
    in count mode, the events at the final timestamp in duration mode): they are not staged."""
    ranges, fo = [], [0]
    for i in img_indices:
        b, e = event_range(i, img_timestamp_event_idx, fixed_duration, nr_events)
        cuts = chunk_cuts(t_events, b, e, nr_events_data, fixed_duration)
        if event_representation == "voxel_grid" and np.any(np.diff(cuts) <= 0):
            # generate_voxel_grid reads events[-1, 2] of every chunk (data_util.py:67): an empty chunk is an IndexError there
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        ranges.append((b, b + int(cuts[-1])))
        fo.extend((fo[-1] + cuts[1:]).tolist())
    t_dev, xyp_dev, _ = stager.stage(t_events, xyp_events, ranges)
    grids = event_tensors(t_dev, xyp_dev, np.asarray(fo, dtype=np.int64), shape, event_representation, nr_temporal_bins,
                          separate_pol, mode)
    return grids.view(len(ranges), -1, shape[0], shape[1])
