"""DDD17 event ingest: the loader mirror and the native-record voxeliser against goldens produced by the
REFERENCE's own load_events / extract_events_from_memmap + the chunking of DDD17Events.__getitem__ + generate_voxel_grid
(oracle/make_golden.py --ddd17).  Integer / float64-replay work: bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import load_golden


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _recording(tmp_path, z):
    z["t"].tofile(os.path.join(tmp_path, "events.dat.t"))
    z["xyp"].tofile(os.path.join(tmp_path, "events.dat.xyp"))
    os.makedirs(os.path.join(tmp_path, "index"), exist_ok=True)
    np.save(os.path.join(tmp_path, "index", "index_50ms.npy"), z["index"])
    return str(tmp_path)


def test_loader_mirror_host_functions_match_reference(tmp_path):
    from openess_b200.datasets.extract_data_tools import example_loader_ddd17 as ld
    z = load_golden("ddd17_ingest")
    d = _recording(tmp_path, z)
    idx, t_ev, xyp_ev, masks = ld.load_files_in_directory(d, 50)
    assert masks == [] and t_ev.shape == (24000, 1) and xyp_ev.shape == (24000, 3) and t_ev.dtype == np.int64
    for s in range(4):
        for fixed in (False, True):
            ev = ld.extract_events_from_memmap(t_ev, xyp_ev, s, idx, fixed, 6000)
            tag = f"s{s}_{int(fixed)}"
            assert ev.dtype == np.int64 and ev.shape == (int(z[tag + "__n"]), 4)
            assert _sha(ev) == str(z[tag + "__sha_events"])
            b, e = ld.event_range(s, idx, fixed, 6000)
            np.testing.assert_array_equal(ld.chunk_cuts(t_ev, b, e, 4, fixed), z[tag + "__cuts"])
    with pytest.raises(IndexError):
        ld.chunk_cuts(t_ev, 5, 5, 4, False)
    # chunk boundaries: count mode tiles the first n - n % chunks records; duration mode is monotone and ends before the
    # records of the final timestamp
    rng = np.random.default_rng(0)
    for _ in range(50):
        b = int(rng.integers(0, 20000))
        e = int(rng.integers(b + 1, 24001))
        nd = int(rng.integers(1, 9))
        c = ld.chunk_cuts(t_ev, b, e, nd, False)
        assert c[0] == 0 and np.all(np.diff(c) == (e - b) // nd) and c[-1] == (e - b) // nd * nd
        d = ld.chunk_cuts(t_ev, b, e, nd, True)
        assert d[0] == 0 and np.all(np.diff(d) >= 0) and d[-1] <= e - b
        ts = t_ev[b:e, 0]
        delta = int((ts[-1] - ts[0]) / nd)
        assert all(int(np.searchsorted(ts, ts[0] + (i + 1) * delta)) == d[i + 1] for i in range(nd))
    with pytest.raises(IndexError):        # fewer records than chunks: an empty chunk is an IndexError in the reference too
        ld.load_event_tensors(None, t_ev, xyp_ev, [3], np.array([[0, 2, 0]] * 4), (260, 346), nr_events_data=4, nr_events=2)
    with pytest.raises(RuntimeError):
        ld.DDD17Stager(16, device="cpu")                       # no CPU path


@pytest.mark.gpu
def test_native_records_bit_exact_vs_reference(tmp_path, oracle):
    from openess_b200 import voxel
    from openess_b200.datasets.extract_data_tools import example_loader_ddd17 as ld
    z = load_golden("ddd17_ingest")
    d = _recording(tmp_path, z)
    idx, t_ev, xyp_ev, _ = ld.load_files_in_directory(d, 50)
    H, W = 260, 346
    stager = ld.DDD17Stager(1000)                              # grows on demand
    for fixed in (False, True):
        out = ld.load_event_tensors(stager, t_ev, xyp_ev, [0, 1, 2, 3], idx, (H, W), nr_events_data=4, nr_events=6000,
                                    fixed_duration=fixed, separate_pol=False)
        assert out.shape == (4, 20, H, W)
        for s in range(4):
            assert _sha(out[s].cpu().numpy()) == str(z[f"s{s}_{int(fixed)}__sha_grid"]), (s, fixed)
    # one frame, both polarity layouts and the histogram: native records == int64 rows through the existing kernels == oracle
    ev = ld.extract_events_from_memmap(t_ev, xyp_ev, 1, idx, False, 6000)
    t_dev = torch.from_numpy(np.ascontiguousarray(ev[:, 2])).cuda()
    xyp_dev = torch.from_numpy(np.ascontiguousarray(ev[:, [0, 1, 3]]).astype(np.int16)).cuda()
    for sep in (False, True):
        a = voxel.voxel_tbilinear_ddd17(t_dev, xyp_dev, 5, H, W, separate_pol=sep)[0].cpu().numpy()
        b = voxel.voxel_tbilinear(torch.from_numpy(ev.copy()).cuda(), 5, H, W, separate_pol=sep)[0].cpu().numpy()
        assert a.tobytes() == b.tobytes() == oracle.voxel_tbilinear(ev.copy(), (H, W), 5, sep).tobytes()
    valid = (ev[:, 0] < W)
    hist = voxel.voxel_histogram_ddd17(t_dev[torch.from_numpy(valid).cuda()].contiguous(),
                                       xyp_dev[torch.from_numpy(valid).cuda()].contiguous(), H, W)[0].cpu().numpy()
    assert hist.tobytes() == oracle.histogram(ev[valid].copy(), (H, W)).tobytes()
    atom = voxel.voxel_tbilinear_ddd17(t_dev, xyp_dev, 5, H, W, separate_pol=False, mode="atomic")[0].cpu().numpy()
    assert np.abs(atom - oracle.voxel_tbilinear(ev.copy(), (H, W), 5, False)).max() < 2e-5
