"""DSEC event ingest (DSEC half): mirror of DSEC/utils/eventslicer.py:10-209 (`EventSlicer`) and of the
record selection of `Sequence.__getitem__` (DSEC/dataset/sequence_ov.py:241-305), feeding the device-side sample assembly.

`EventSlicer` keeps the reference's constructor and methods (`get_events`, `get_events_fixed_num`,
`get_events_fixed_num_recurrent`, `get_conservative_window_ms`, `get_conservative_ms`, `get_time_indices_offsets`, `ms2idx`,
`get_start_time_us`, `get_final_time_us`) and return values.  `h5f` is anything indexable like the reference's h5py.File
(`h5f['events/x']`, `h5f['ms_to_idx']`, optional `h5f['t_offset'][()]`): an open h5py file where h5py / hdf5plugin exist, a
dict of numpy arrays / memory maps otherwise -- this module imports neither h5py nor numba (the reference's numba loops over
a sorted millisecond window are two `np.searchsorted` calls).

B200 wire format: the index ranges are resolved on the host (`fixed_num_range`, `window_range`: integer searches over the
1 kHz index), the records themselves are copied ONCE from the file into pinned host buffers as they are on disk (x, y
uint16, t uint32, p uint8: 9 B / event instead of the 32 B / event float64 [n, 4] array `np.stack([x_rect, y_rect, t, p])`
builds at sequence_ov.py:301) and uploaded asynchronously; rectification, time normalisation and voxelisation of all
`nr_events_data` chunks of all samples of a batch happen on the device in one call
(`openess_b200.voxel.dsec_events_to_voxel_grid` / `OpenESSPretrainStep.event_tensor` on the `RawEvents` this returns).
"""
import math

import numpy as np
import torch


class EventSlicer:
    """DSEC/utils/eventslicer.py:10-209."""

    def __init__(self, h5f):
        self.h5f = h5f
        self.events = dict()
        for dset_str in ['p', 'x', 'y', 't']:
            self.events[dset_str] = self.h5f['events/{}'.format(dset_str)]
        self.ms_to_idx = np.asarray(self.h5f['ms_to_idx'], dtype='int64')
        if "t_offset" in list(h5f.keys()):
            self.t_offset = int(h5f['t_offset'][()])
        else:
            self.t_offset = 0
        self.t_final = int(self.events['t'][-1]) + self.t_offset

    def get_start_time_us(self):
        return self.t_offset

    def get_final_time_us(self):
        return self.t_final

    # ---- index resolution (host integers only) ------------------------------------------------------------------------
    def window_range(self, t_start_us, t_end_us):
        """[begin, end) record range of get_events(t_start_us, t_end_us) (:45-58), or None outside the 1 kHz index."""
        assert t_start_us < t_end_us
        t_start_us -= self.t_offset
        t_end_us -= self.t_offset
        t_start_ms, t_end_ms = self.get_conservative_window_ms(t_start_us, t_end_us)
        a, b = self.ms2idx(t_start_ms), self.ms2idx(t_end_ms)
        if a is None or b is None:
            return None
        tw = np.asarray(self.events['t'][a:b])
        i0, i1 = self.get_time_indices_offsets(tw, t_start_us, t_end_us)
        return int(a + i0), int(a + i1)

    def fixed_num_range(self, t_end_us, nr_events=100000):
        """[begin, end) record range of get_events_fixed_num(t_end_us, nr_events) (:78-94), or None."""
        t_end_us -= self.t_offset
        lo_ms, hi_ms = self.get_conservative_ms(t_end_us)
        a, b = self.ms2idx(lo_ms), self.ms2idx(hi_ms)
        if a is None or b is None:
            return None
        tw = np.asarray(self.events['t'][a:b])
        _, i1 = self.get_time_indices_offsets(tw, t_end_us, t_end_us)
        end = int(a + i1)
        return max(end - nr_events, 0), end

    # ---- reference methods --------------------------------------------------------------------------------------------
    def get_events(self, t_start_us, t_end_us, max_events_per_data=-1):
        r = self.window_range(t_start_us, t_end_us)
        if r is None:
            print('Error', 'start', t_start_us - self.t_offset, 'end', t_end_us - self.t_offset)
            return None
        b, e = r
        events = {'t': np.asarray(self.events['t'][b:e]) + self.t_offset}
        for dset_str in ['p', 'x', 'y']:
            events[dset_str] = np.asarray(self.events[dset_str][b:e])
            assert events[dset_str].size == events['t'].size
        return events

    def get_events_fixed_num(self, t_end_us, nr_events=100000):
        r = self.fixed_num_range(t_end_us, nr_events)
        if r is None:
            return None
        return self.get_events_fixed_num_recurrent(*r) if r[0] < r[1] else \
            {k: np.asarray(v[r[0]:r[1]]) for k, v in self.events.items()}

    def get_events_fixed_num_recurrent(self, t_start_us_idx, t_end_us_idx):
        assert t_start_us_idx < t_end_us_idx
        return {k: np.asarray(v[t_start_us_idx:t_end_us_idx]) for k, v in self.events.items()}

    @staticmethod
    def get_conservative_window_ms(ts_start_us, ts_end_us):
        assert ts_end_us > ts_start_us
        return math.floor(ts_start_us / 1000), math.ceil(ts_end_us / 1000)

    @staticmethod
    def get_conservative_ms(ts_us):
        return math.floor(ts_us / 1000), math.ceil(ts_us / 1000)

    @staticmethod
    def get_time_indices_offsets(time_array, time_start_us, time_end_us):
        """:152-203 on a time-sorted window: first index with t >= start, first index with t >= end."""
        assert time_array.ndim == 1
        if time_array.size == 0:
            # a query time exactly on a millisecond gives an empty conservative window (floor == ceil).  The reference's
            # numba loop then reads time_array[-1] out of bounds (nopython mode does not check); its only non-asserting
            # outcome is (size, size) = (0, 0), which is also the meaningful answer: the boundary is the window start.
            return 0, 0
        if time_array[-1] < time_start_us:
            return time_array.size, time_array.size
        return (int(np.searchsorted(time_array, time_start_us, side='left')),
                int(np.searchsorted(time_array, time_end_us, side='left')))

    def ms2idx(self, time_ms):
        assert time_ms >= 0
        if time_ms >= self.ms_to_idx.size:
            return None
        return self.ms_to_idx[time_ms]


def sample_chunks(slicer, ts_end, nr_events_data=20, nr_events_per_data=100000, fixed_duration=False, delta_t_us=None):
    """Record ranges of the `nr_events_data` chunks of one sample, as Sequence.__getitem__ selects them
    (sequence_ov.py:247-254 duration mode, :282-305 count mode) -> list of (begin, end) or None if a lookup fails.
    Count mode: the last `nr_events` records before ts_end cut into equal chunks, the remainder dropped (:212-215)."""
    if fixed_duration:
        ts_start = ts_end - delta_t_us
        per = delta_t_us / nr_events_data
        out = []
        for i in range(nr_events_data):
            r = slicer.window_range(ts_start + i * per, ts_start + (i + 1) * per)
            if r is None:
                return None
            out.append(r)
        return out
    r = slicer.fixed_num_range(ts_end, nr_events_data * nr_events_per_data)
    if r is None:
        return None
    b, e = r
    per = (e - b) // nr_events_data
    return [(b + i * per, b + (i + 1) * per) for i in range(nr_events_data)]


class DSECStager:
    """Pinned host staging of raw DSEC records for a batch of samples + one asynchronous upload per field."""

    FIELDS = (("x", torch.uint16), ("y", torch.uint16), ("t", torch.uint32), ("p", torch.uint8))

    def __init__(self, capacity_events, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DSECStager stages records for the CUDA voxeliser (no CPU path)")
        self._done = None
        self._alloc(int(capacity_events))

    def _alloc(self, cap):
        self.capacity = cap
        self.host = {k: torch.empty(cap, dtype=dt).pin_memory() for k, dt in self.FIELDS}
        self.dev = {k: torch.empty(cap, dtype=dt, device=self.device) for k, dt in self.FIELDS}

    def stage(self, slicer, ranges):
        """ranges: [(begin, end), ...], one per FRAME (chunk), in batch order.  Returns (x, y, t, p device views,
        frame_offsets int64 [F + 1] host tensor); valid until the next stage() on this stager."""
        total = sum(e - b for b, e in ranges)
        if total > self.capacity:
            torch.cuda.current_stream(self.device).synchronize()
            self._alloc(max(total, 2 * self.capacity))
        elif self._done is not None:
            self._done.synchronize()
        offs = [0]
        views = {k: self.host[k].numpy() for k, _ in self.FIELDS}
        for b, e in ranges:
            o = offs[-1]
            for k, _ in self.FIELDS:
                views[k][o:o + e - b] = slicer.events[k][b:e]
            offs.append(o + e - b)
        with torch.cuda.device(self.device):
            for k, _ in self.FIELDS:
                self.dev[k][:total].copy_(self.host[k][:total], non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record()
        return tuple(self.dev[k][:total] for k, _ in self.FIELDS) + (torch.tensor(offs, dtype=torch.int64),)


def stage_raw_events(stager, slicer, timestamps, rectify_map, nr_events_data=20, nr_events_per_data=100000,
                     fixed_duration=False, delta_t_us=None, crop_h=None):
    """Batch of samples (label timestamps of ONE sequence) -> `RawEvents` for OpenESSPretrainStep / dsec_events_to_voxel_grid."""
    from ...training.pretrain_step import RawEvents
    ranges = []
    for ts in timestamps:
        ch = sample_chunks(slicer, int(ts), nr_events_data, nr_events_per_data, fixed_duration, delta_t_us)
        if ch is None:
            raise IndexError(f"timestamp {ts} is outside the recording's millisecond index")
        ranges.extend(ch)
    x, y, t, p, fo = stager.stage(slicer, ranges)
    Hs, Ws = rectify_map.shape[:2]
    return RawEvents(x, y, t, p, fo, rectify_map, (Hs, Ws), Hs - 40 if crop_h is None else crop_h)
