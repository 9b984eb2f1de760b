"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200 import losses, ops, voxel  # noqa: E402

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
# strip splat: dense rows (window path with cuts, runs >= 32 -> sweep path), several strips, odd width (scalar write-out)
for (Hs, Ws, ns) in ((24, 200, 40000), (16, 70, 9000)):
    xs = torch.from_numpy(rng.uniform(-1.2, Ws + 0.2, ns).astype(np.float32)).to(dev)
    ys = torch.from_numpy(rng.uniform(-1.2, Hs + 0.2, ns).astype(np.float32)).to(dev)
    xs[: ns // 4] = xs[: ns // 4].round() % 3 + 65.4       # a hot cell column crossing a strip boundary
    ps = torch.from_numpy(rng.integers(0, 2, ns).astype(np.float32)).to(dev)
    ts = torch.from_numpy(np.sort(rng.random(ns)).astype(np.float32)).to(dev)
    voxel.voxel_trilinear(xs, ys, ps, ts, C, Hs, Ws, mode="ordered")
    voxel.voxel_trilinear(xs, ys, ps, ts.flip(0).contiguous(), C, Hs, Ws, mode="ordered")     # robust path, long ranges
losses.argmax_confusion(lg.detach(), tg, 255)
# tensor-core kernels (tcgen05 / TMA / TMEM), BatchNorm, fused teacher tail
a = torch.randn(300, 96, device=dev)
b = torch.randn(200, 96, device=dev)
ops.gemm_tf32(a, b, torch.randn(200, device=dev))
Cc = 64
wgt = torch.randn(4 * Cc, 2 * Cc, 3, 3, device=dev) * 0.03
wp, bp = ops.convlstm_pack(wgt, torch.randn(4 * Cc, device=dev) * 0.1, Cc)
xc = torch.randn(2, Cc, 13, 22, device=dev)
h1, c1 = ops.convlstm_step(xc, None, wp, bp)
ops.convlstm_step(xc, (h1, c1), wp, bp)
wc = torch.randn(48, 32, 5, 5, device=dev) * 0.05
yc = ops.conv2d_tc(torch.randn(2, 32, 17, 23, device=dev), ops.conv2d_pack(wc), torch.randn(48, device=dev), 5, 2, 2, 1, relu=True)
bn = torch.nn.BatchNorm2d(64).to(dev).train()
w3 = torch.randn(64, 32, 3, 3, device=dev) * 0.05
xb = torch.randn(2, 32, 11, 19, device=dev)
ops.conv_bn_train(xb, ops.conv2d_pack(w3), None, 3, 1, 2, 2, bn, residual=torch.randn(2, 64, 11, 19, device=dev), relu=True)
ops.batchnorm_nhwc_(ops.conv2d_tc(xb, ops.conv2d_pack(w3), None, 3, 1, 1, 1), bn.eval(), relu=True)
ops.planes_to_nhwc_padded(torch.randn(2, 5, 9, 14, device=dev), 8)
d = torch.randn(2, 256, 6, 9, device=dev, requires_grad=True)
sp = torch.randint(0, 7, (2, 24, 36), device=dev)
ops.upnorm_pool(d, sp, 7, 14).square().sum().backward()
# rows added in the fourth session: ViT kernels (incl. the tcgen05 attention with ragged tails), GEMM epilogues, pooling,
# the 7x7 stem through the padded repack, DDD17 native records
qkv = torch.randn(2 * 150, 3 * 128, device=dev)
ops.mha_fwd(qkv, 2, 150, 2, tensor_cores=True)
ops.mha_fwd(qkv, 2, 150, 2, tensor_cores=False)
ops.mha_fwd(qkv[:7].contiguous(), 1, 7, 2, tensor_cores=True)
xr = torch.randn(2 * 36, 128, device=dev)
ops.layernorm_rows(xr, torch.ones(128, device=dev), torch.zeros(128, device=dev), 1e-6)
ops.l2norm_rows_(xr.clone())
rows_, hw_ = ops.vit_patchify(torch.rand(2, 3, 72, 100, device=dev), 16)
ops.vit_assemble(torch.randn(2 * 35, 128, device=dev), torch.randn(128, device=dev), torch.randn(36, 128, device=dev), 2, 36)
ops.bilinear_tokens_to_nchw(torch.randn(2 * 35, 11, device=dev), 2, 5, 7, (72, 100))
res_ = torch.randn(300, 200, device=dev)
ops.gemm_tf32_ex(a, b, torch.randn(200, device=dev), residual=res_, act="gelu", out=res_, round_out=True)
ops.gemm_tf32_ex(a, torch.randn(11, 96, device=dev))                       # N = 11: scalar epilogue tail
xm = torch.randn(2, 64, 13, 17, device=dev).contiguous(memory_format=torch.channels_last)
ops.maxpool3x3s2_nhwc(xm)
ops.global_avgpool_nhwc(xm)
w7 = torch.zeros(64, 8, 7, 7, device=dev)
w7[:, :3] = torch.randn(64, 3, 7, 7, device=dev) * 0.05
ops.conv2d_tc(ops.planes_to_nhwc_padded(torch.rand(2, 3, 37, 45, device=dev), 8), ops.conv2d_pack(w7), None, 7, 2, 3, 1, relu=True,
              round_out=True)
nd = 5000
td = torch.from_numpy(np.sort(rng.integers(0, 50000, nd)).astype(np.int64) + 1_500_000_000).to(dev)
xypd = torch.from_numpy(np.stack([rng.integers(0, 346, nd), rng.integers(0, 260, nd), rng.integers(0, 2, nd)], 1).astype(np.int16)).to(dev)
fod = torch.tensor([0, 1200, 1200, nd], dtype=torch.int64)
voxel.voxel_tbilinear_ddd17(td, xypd, 5, 260, 346, frame_offsets=fod, separate_pol=False)
voxel.voxel_tbilinear_ddd17(td, xypd, 5, 260, 346, frame_offsets=fod, separate_pol=True, mode="atomic")
voxel.voxel_histogram_ddd17(td, xypd, 260, 346, frame_offsets=fod)
# rows added in round 2: thin-input head conv through the overlapping-window tensor map, reconstruction post-processing,
# InfoNCE on the tensor cores (3xTF32 operand splitting; OESS_INFONCE=tc forces it for a small, ragged M)
w5 = torch.zeros(32, 8, 5, 5, device=dev)
w5[:, :5] = torch.randn(32, 5, 5, 5, device=dev) * 0.05
x5 = ops.planes_to_nhwc_padded_w(torch.randn(2, 5, 19, 37, device=dev), 8, 2)
ops.conv2d_rowunfold(x5, ops.conv2d_pack_rowunfold(w5), torch.randn(32, device=dev), 5, 5, 37, relu=True)
from openess_b200.e2vid.image_reconstructor import gaussian_kernel_5x5  # noqa: E402
ops.unsharp_rescale(torch.rand(2, 1, 21, 33, device=dev), gaussian_kernel_5x5(1.0).to(dev), 0.3, 0.0, 1.0)
if os.environ.get("OESS_INFONCE", "")[:1] == "t":
    kk = torch.nn.functional.normalize(torch.randn(133, 64, device=dev), dim=1).requires_grad_(True)
    qq = torch.nn.functional.normalize(torch.randn(133, 64, device=dev), dim=1).requires_grad_(True)
    losses.infonce(kk, qq, 0.07).backward()
# persistent tensor-core kernels (second session of round 2) at sizes where every CTA walks SEVERAL tiles: operand ring across
# tile boundaries, both TMEM accumulators, staging-block reuse under TMA stores, statistics flush on a Cout-tile change
a2 = torch.randn(128 * 450 + 5, 64, device=dev)
r2 = torch.randn(a2.shape[0], 256, device=dev)
ops.gemm_tf32_ex(a2, torch.randn(256, 64, device=dev), torch.randn(256, device=dev), residual=r2, out=r2)
ops.gemm_tf32(a2[:19000], torch.randn(72, 64, device=dev))                      # BN = 128, ragged last N tile
xl = torch.randn(2, 64, 120, 160, device=dev).contiguous(memory_format=torch.channels_last)
w1 = torch.randn(512, 64, 1, 1, device=dev) * 0.1
bn2 = torch.nn.BatchNorm2d(512).to(dev).train()
ops.conv_bn_train(xl, ops.conv2d_pack(w1), None, 1, 1, 0, 1, bn2, relu=True)     # 300 pixel tiles x 2 Cout tiles, fused statistics
ops.conv2d_tc(xl, ops.conv2d_pack(torch.randn(64, 64, 3, 3, device=dev) * 0.05), None, 3, 2, 1, 1, relu=True, residual=None)
xlb = xl.to(torch.bfloat16)
ops.conv2d_tc_bf16(xlb, ops.conv2d_pack_bf16(w1), None, 1, 1, 0, 1, relu=True)
ops.conv_bn_train_bf16(xlb, ops.conv2d_pack_bf16(w1), None, 1, 1, 0, 1, bn2, relu=True, want_f32=False)
wl = torch.randn(4 * 64, 128, 3, 3, device=dev) * 0.05
wpl, bpl = ops.convlstm_pack(wl, torch.randn(256, device=dev) * 0.1, 64)
hl, cl_ = ops.convlstm_step(xl, None, wpl, bpl)                                  # 150 tiles x 2 samples
ops.convlstm_step(xl, (hl, cl_), wpl, bpl)
hb, cb = ops.convlstm_step_bf16(xlb, None, wpl.to(torch.bfloat16), bpl)
ops.convlstm_step_bf16(xlb, (hb, cb), wpl.to(torch.bfloat16), bpl)
# CTA-pair kernels (cta_group::2): remote barrier arrives, 2-SM TMA loads and commits, a dummy peer tile, ragged N / Cout tiles
a3 = torch.randn(256 * 80 + 100, 96, device=dev)
ops.gemm_tf32_ex(a3, torch.randn(520, 96, device=dev), torch.randn(520, device=dev))    # 192-wide pair tiles, ragged last N tile
ops.gemm_tf32_ex(a3, torch.randn(256, 96, device=dev), torch.randn(256, device=dev))    # 256-wide pair tiles
xp = torch.randn(3, 128, 40, 72, device=dev).contiguous(memory_format=torch.channels_last)
wp3 = torch.randn(320, 128, 3, 3, device=dev) * 0.03
bn3 = torch.nn.BatchNorm2d(256).to(dev).train()
ops.conv_bn_train(xp, ops.conv2d_pack(wp3[:256].contiguous()), None, 3, 1, 1, 1, bn3, relu=True)   # fused statistics, one Cout tile
ops.conv2d_tc(xp, ops.conv2d_pack(wp3), torch.randn(320, device=dev), 3, 1, 1, 1, relu=True)       # ragged second Cout tile
ops.conv2d_tc_bf16(xp.to(torch.bfloat16), ops.conv2d_pack_bf16(wp3), None, 3, 1, 1, 1, relu=True)
xu = torch.randn(2, 32, 9, 11, device=dev, requires_grad=True)
su = torch.randn(2, 16, 18, 22, device=dev, requires_grad=True)
ops.upsample2x_cat(xu, su).square().sum().backward()
xr = torch.randn(2, 5, 7, 9, device=dev, requires_grad=True)
ops.bilinear_resize(xr, (30, 23)).square().sum().backward()
torch.cuda.synchronize()
print("sanitize smoke done")
