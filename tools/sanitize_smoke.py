"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import losses, voxel  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
C, H, W = 5, 48, 64
sizes = [3000, 0, 1, 5000, 700]
n = sum(sizes)
fo = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
x = torch.from_numpy(rng.uniform(-1.2, W + 0.2, n).astype(np.float32)).to(dev)
y = torch.from_numpy(rng.uniform(-1.2, H + 0.2, n).astype(np.float32)).to(dev)
x[:2000] = x[:2000].round() % 6 + 10.3          # hot cells: long runs, same-accumulator rounds
y[:2000] = y[:2000].round() % 4 + 7.6
p = torch.from_numpy(rng.integers(0, 2, n).astype(np.float32)).to(dev)
t = torch.from_numpy(np.concatenate([np.sort(rng.random(s)) for s in sizes]).astype(np.float32)).to(dev)
for mode in ("ordered", "atomic"):
    o = voxel.voxel_trilinear(x, y, p, t, C, H, W, frame_offsets=fo, mode=mode, normalize=True)
t2 = t.clone()
t2[:3000] = t2[:3000].flip(0)                     # unsorted time -> robust (match_any) path
voxel.voxel_trilinear(x, y, p, t2, C, H, W, frame_offsets=fo, mode="ordered")
voxel.voxel_trilinear(x[:500], y[:500], p[:500], t[:500].sort().values, 3, 1030, 12, mode="ordered")   # generic path
ev = torch.stack([x.clamp(0, W - 1).long(), y.clamp(0, H - 1).long(), (t * 50000).long(), p.long()], 1).contiguous()
for mode in ("ordered", "atomic"):
    voxel.voxel_tbilinear(ev, C, H, W, frame_offsets=fo, separate_pol=True, mode=mode)
    voxel.voxel_tbilinear(ev.double(), 7, H, W, frame_offsets=fo, separate_pol=False, mode=mode)
voxel.voxel_histogram(ev, H, W, frame_offsets=fo)
rmap = torch.rand(H, W, 2, device=dev) * W
voxel.dsec_events_to_voxel_grid(ev[:, 0].to(torch.uint16), ev[:, 1].to(torch.uint16), ev[:, 2].to(torch.uint32),
                                ev[:, 3].to(torch.uint8), rmap, C, frame_offsets=fo)
feat = torch.randn(2, 32, H, W, device=dev, requires_grad=True)
seg = torch.randint(0, 10, (2, H, W), device=dev)
k = losses.superpixel_pool(feat, seg, 10)
q = losses.superpixel_pool(torch.randn(2, 32, H, W, device=dev), seg, 10)
losses.infonce(k, q, 0.07).backward()
lg = torch.randn(2, 11, H, W, device=dev, requires_grad=True)
tg = torch.randint(0, 11, (2, H, W), device=dev)
tg[0, :3] = 255
(losses.dice_ce(lg, tg, 255) + losses.cosine_consistency(lg, lg.detach() * 0.5 + 1) + losses.l1_mean(lg, lg.detach() + 1)).backward()
losses.confusion(lg.argmax(1), tg, 11, 255)
losses.convlstm_gates(torch.randn(2, 16, 6, 8, device=dev), torch.randn(2, 4, 6, 8, device=dev))
torch.cuda.synchronize()
print("sanitize smoke done")
