"""DSEC event ingest (DSEC half): the EventSlicer mirror against goldens produced by the REFERENCE's own
DSEC/utils/eventslicer.py (numba search loops) on a synthetic recording (oracle/make_golden.py --dsec-slicer), and the staged
`RawEvents` path against the oracle voxeliser.  Index / byte work: exact."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.make_golden import synth_dsec_recording


def _sha(ev):
    a = np.concatenate([ev[k].astype(np.int64) for k in ("x", "y", "t", "p")])
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_eventslicer_mirror_matches_reference():
    from openess_b200.DSEC.utils.eventslicer import EventSlicer, sample_chunks
    z = load_golden("dsec_slicer")
    rec = synth_dsec_recording()
    sl = EventSlicer(rec)
    assert sl.get_start_time_us() == int(z["start"]) and sl.get_final_time_us() == int(z["final"])
    for i in range(7):
        a, b = (int(v) for v in z[f"w{i}__q"])
        ev = sl.get_events(a, b)
        assert (ev is None) == bool(z[f"w{i}__none"]), i
        if ev is not None:
            assert ev["t"].size == int(z[f"w{i}__n"]) and _sha(ev) == str(z[f"w{i}__sha"]), i
            assert ev["x"].dtype == np.uint16 and ev["p"].dtype == np.uint8   # t: `t_window + t_offset`, numpy's promotion as in the reference
    for i in range(7):
        te, n = (int(v) for v in z[f"f{i}__q"])
        ev = sl.get_events_fixed_num(te, n)
        assert (ev is None) == bool(z[f"f{i}__none"]), i
        if ev is not None:
            assert ev["t"].size == int(z[f"f{i}__n"]) and _sha(ev) == str(z[f"f{i}__sha"]), i
            assert ev["t"].dtype == np.uint32                       # no t_offset on this path, as the reference
            ch = sample_chunks(sl, te, nr_events_data=4, nr_events_per_data=n // 4 if n % 4 == 0 else n)
            if n % 4 == 0:
                per = int(z[f"f{i}__per"])
                assert [e - b for b, e in ch] == [per] * 4
                if per:
                    assert [int(rec["events/t"][b]) for b, _ in ch] == z[f"f{i}__chunk_t0"].tolist()
    ts_end, delta, nd = (int(v) for v in z["d__q"])
    ch = sample_chunks(sl, ts_end, nr_events_data=nd, fixed_duration=True, delta_t_us=delta)
    assert [e - b for b, e in ch] == z["d__n"].tolist()
    for i, (b, e) in enumerate(ch):
        ev = {k: np.asarray(rec["events/" + k][b:e]) for k in "xyp"}
        ev["t"] = np.asarray(rec["events/t"][b:e]) + int(rec["t_offset"])
        assert _sha(ev) == str(z[f"d{i}__sha"])
    assert sample_chunks(sl, int(rec["t_offset"]) + 700_001, 4, 1000) is None
    from openess_b200.DSEC.utils import eventslicer as m
    with pytest.raises(RuntimeError):
        m.DSECStager(16, device="cpu")                               # no CPU path


def test_eventslicer_ranges_satisfy_their_definition():
    """Size-independent property on random recordings: window_range is exactly {i : t_start <= t_i < t_end} inside the 1 kHz
    index, fixed_num_range ends at the first record with t >= t_end and holds min(n, available) records, chunks tile it."""
    from openess_b200.DSEC.utils.eventslicer import EventSlicer, sample_chunks
    rng = np.random.default_rng(7)
    for trial in range(20):
        n, span_ms = int(rng.integers(50, 4000)), int(rng.integers(20, 200))
        t = np.sort(rng.integers(0, span_ms * 1000, n)).astype(np.uint32)
        if trial % 3 == 0:
            t[n // 2: n // 2 + 30] = t[n // 2]                      # bursts on one microsecond
            t = np.sort(t)
        off = int(rng.integers(0, 10_000_000))
        rec = {"events/x": np.zeros(n, np.uint16), "events/y": np.zeros(n, np.uint16), "events/t": t, "events/p": np.zeros(n, np.uint8),
               "ms_to_idx": np.searchsorted(t, np.arange(span_ms + 1, dtype=np.int64) * 1000, side="left").astype(np.uint64),
               "t_offset": np.array(off)}
        sl = EventSlicer(rec)
        for _ in range(30):
            a = int(rng.integers(0, span_ms * 1000 - 1))
            b = int(rng.integers(a + 1, span_ms * 1000 + 1))
            r = sl.window_range(off + a, off + b)
            assert r is not None
            idx = np.nonzero((t >= a) & (t < b))[0]
            assert (r[1] - r[0]) == idx.size and (idx.size == 0 or (r[0] == idx[0] and r[1] == idx[-1] + 1))
            te, k = int(rng.integers(0, span_ms * 1000 + 1)), int(rng.integers(1, n + 50))
            r = sl.fixed_num_range(off + te, k)
            end = int(np.searchsorted(t, te, side="left"))
            assert r == (max(end - k, 0), end)
            ch = sample_chunks(sl, off + te, nr_events_data=4, nr_events_per_data=max(k // 4, 1))
            per = (ch[0][1] - ch[0][0])
            assert all(e - b_ == per for b_, e in ch) and all(ch[i][1] == ch[i + 1][0] for i in range(3)) and ch[-1][1] <= end
        assert sl.window_range(off + span_ms * 1000, off + span_ms * 1000 + 2000) is None      # past the index


@pytest.mark.gpu
def test_staged_raw_events_voxelise_like_the_reference_path(oracle):
    """Two samples x 4 chunks staged from the recording -> RawEvents -> device voxel grids == oracle on the host arrays the
    reference would have built (rectify_events + events_to_voxel_grid + VoxelGrid.convert per chunk)."""
    from openess_b200 import voxel
    from openess_b200.DSEC.utils.eventslicer import DSECStager, EventSlicer, stage_raw_events
    rec = synth_dsec_recording()
    sl = EventSlicer(rec)
    off = int(rec["t_offset"])
    rng = np.random.default_rng(3)
    rmap = np.stack(np.meshgrid(np.arange(640, dtype=np.float32), np.arange(480, dtype=np.float32)), -1)
    rmap = (rmap + rng.uniform(-0.75, 0.75, rmap.shape)).astype(np.float32)
    rmap_d = torch.from_numpy(rmap).cuda()
    stager = DSECStager(1000)
    raw = stage_raw_events(stager, sl, [off + 400_000, off + 650_000], rmap_d, nr_events_data=4, nr_events_per_data=5000)
    assert raw.frame_offsets.tolist() == list(range(0, 40001, 5000)) and raw.crop_h == 440
    grids = voxel.dsec_events_to_voxel_grid(raw.x, raw.y, raw.t, raw.p, rmap_d, 5, frame_offsets=raw.frame_offsets.cuda(),
                                            mode="ordered").cpu().numpy()
    f = 0
    for ts in (off + 400_000, off + 650_000):
        ev = sl.get_events_fixed_num(ts, 20000)
        for i in range(4):
            s = slice(i * 5000, (i + 1) * 5000)
            x, y, pol, t = oracle.dsec_rectify_tnorm(ev["x"][s], ev["y"][s], ev["t"][s], ev["p"][s], rmap)
            assert grids[f].tobytes() == oracle.voxel_trilinear(x, y, pol, t, 5, 480, 640).tobytes(), f
            f += 1
