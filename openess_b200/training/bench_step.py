"""Measurement of the end-to-end OpenESS pretraining step (BASELINE config 3 per-GPU shard), used by
bench.py (`train_step` block of the bench line) and tools/bench_train_step.py.

Ours:   raw DSEC records (20 x 100 000 events / sample, pinned host memory) -> H2D -> rectify + voxelise (bit-exact ordered mode)
        -> EventPreprocessor + E2VID x 20 (tcgen05) -> SemSegE2VID ; frame -> dilated ResNet-50 teacher (tcgen05, train-mode BN)
        -> fused superpixel pooling -> InfoNCE + Dice/CE -> backward -> bucketed NCCL gradient all-reduce overlapped with
        backward (world > 1) -> 2 x AdamW -> D2H of the loss.
Literal baseline (same GPU, same weights): the torch / cuDNN formulation of training/pretrain_trainer.py:427-472 + :550-562 +
        utils/loss_functions.py (dense [B, 100, 440, 640] event tensor from pinned host memory as the reference's loader hands it
        over, cuDNN convolutions with torch's default TF32 setting, sparse one-hot pooling on permuted copies, softmax / one-hot
        Dice + CE, cuBLAS InfoNCE), none of this repository's kernels on its path.
Weights are seeded random (no checkpoints offline); inputs synthetic (openess_b200/utils/synth.py)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

E2VID_CFG = {'num_bins': 5, 'skip_type': 'sum', 'recurrent_block_type': 'convlstm', 'num_encoders': 3,
             'base_num_channels': 32, 'num_residual_blocks': 2, 'norm': 'BN', 'use_upsample_conv': False}


def _seeded_state_dict():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    tests = os.path.join(root, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    from seeded_weights import seeded_state_dict
    return seeded_state_dict


def build_modules(dev, K=11):
    from ..e2vid.model.model import E2VIDRecurrent
    from ..models.image_model import DilationFeatureExtractor
    from ..models.style_networks import SemSegE2VID
    ssd = _seeded_state_dict()
    e2vid = E2VIDRecurrent(E2VID_CFG, latent_only=True)
    e2vid.load_state_dict(ssd(e2vid, 1205), strict=True)
    e2vid = e2vid.eval().to(dev).fold_bn()
    torch.manual_seed(1205)
    back = SemSegE2VID(input_c=256, output_c=K, skip_connect=True, skip_type='concat', text_embeddings_path='').to(dev)
    teacher = DilationFeatureExtractor()
    teacher.load_state_dict(ssd(teacher, 77), strict=True)
    return e2vid, back, teacher.to(dev)


def literal_step(e2vid, back, teacher, event, frame, pl, sp, S, steps):
    """pretrain_trainer.py:427-472 with the reference's own torch formulation of every block."""
    feat_frame = teacher(frame)                                                     # :434
    states = None
    for i in range(steps):                                                          # :437-441
        ev = event[:, 5 * i:5 * i + 5]
        nz = ev != 0                                                                # inference_utils.py:77-85
        n = nz.sum()
        mean = ev.sum() / n
        std = torch.sqrt((ev ** 2).sum() / n - mean ** 2)
        ev = nz.float() * (ev - mean) / std
        with torch.no_grad():
            _, states, latent = e2vid(ev, states)
    pred, feat_voxel = back({k: v.detach() for k, v in latent.items()})            # :551-553
    logits = pred[1]
    ce = F.cross_entropy(logits, pl, ignore_index=255)                              # loss_functions.py:17-24
    mask = (pl != 255)
    onehot = F.one_hot((pl * mask).long(), logits.shape[1]).permute(0, 3, 1, 2).float() * mask[:, None]
    prob = logits.softmax(1) * mask[:, None]
    dice = 0
    for c in range(logits.shape[1]):                                                # loss_functions.py:80-90, 114-135
        num = 2 * (prob[:, c] * onehot[:, c]).sum() + 1
        den = (prob[:, c] ** 2 + onehot[:, c] ** 2).sum() + 1
        dice = dice + (1 - num / den)
    loss_dense = dice / logits.shape[1] + ce
    B = feat_voxel.shape[0]
    spx = torch.arange(0, B * S, S, device=sp.device)[:, None, None] + sp          # :446-449
    sI = spx.flatten()
    idx = torch.arange(sI.shape[0], device=sp.device)
    with torch.no_grad():
        one_hot = torch.sparse_coo_tensor(torch.stack((sI, idx), 0), torch.ones(sI.shape[0], device=sp.device))
    cnt = torch.sparse.sum(one_hot, 1).to_dense()[:, None] + 1e-6
    k = (one_hot @ feat_voxel.permute(0, 2, 3, 1).flatten(0, 2)) / cnt             # :456-459
    q = (one_hot @ feat_frame.permute(0, 2, 3, 1).flatten(0, 2)) / cnt             # :461-463
    nce = F.cross_entropy((k @ q.t()) / 0.07, torch.arange(k.shape[0], device=k.device))    # loss_functions.py:147-153
    return nce + loss_dense, nce, loss_dense


def run(batch=4, steps=5, warmup=2, events=100_000, baseline_steps=0, rank=0, local=0, world=1):
    """Returns a dict (rank 0; None elsewhere).  Under torchrun (world > 1) every rank runs its own shard and the gradients
    are all-reduced over NCCL; times are CUDA-event times, max over ranks."""
    import torch.distributed as dist
    from .. import _lib
    from ..e2vid.image_reconstructor import ImageReconstructor
    from ..e2vid.model import model as e2vid_model
    from ..models import image_model as im
    from ..models import style_networks as sn
    from ..utils import synth
    from ..utils.loss_functions import NCELoss, TaskLoss
    from .pretrain_step import OpenESSPretrainStep, RawEvents

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = True                # torch default: what the reference's own GPU run uses
    B, Hs, Ws, Hc, K, S, NF = batch, 480, 640, 440, 11, 100, 20
    e2vid, back, teacher = build_modules(dev, K)
    opts = SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False, no_recurrent=False)
    rec = ImageReconstructor(e2vid, Hc, Ws, 5, dev, opts)
    step = OpenESSPretrainStep(rec, back, teacher, TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255),
                               NCELoss(temperature=0.07), nr_events_data_b=NF, superpixel_size=S, data_parallel=world > 1)

    rng = np.random.default_rng(1205 + rank)
    rmap = torch.from_numpy(synth.synth_rectify_map(rng)).to(dev)
    x, y, t, p = synth.synth_raw_frames(rng, B * NF, n=events)
    pin = lambda a: torch.from_numpy(a).pin_memory()      # noqa: E731
    ev = RawEvents(pin(x), pin(y), pin(t), pin(p), torch.arange(0, (B * NF + 1) * events, events, dtype=torch.int64),
                   rmap, (Hs, Ws), Hc)
    frame = torch.from_numpy(rng.random((B, 3, Hc, Ws)).astype(np.float32)).pin_memory()
    pl = rng.integers(0, K, (B, Hc, Ws))
    pl[rng.random(pl.shape) < 0.02] = 255
    pl = torch.from_numpy(pl.astype(np.int64)).pin_memory()
    sp = torch.from_numpy(synth.synth_superpixels(rng, B, Hc, Ws, S)).pin_memory()
    data = (ev, None, frame, pl, sp)
    h2d = int(x.nbytes + y.nbytes + t.nbytes + p.nbytes + frame.numel() * 4 + pl.numel() * 8 + sp.numel() * 8)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # every step copies its whole batch from pinned host memory; the copy of step i + 1 runs on a side stream under step i
    from .prefetch import DevicePrefetcher
    prefetch = os.environ.get("OESS_PREFETCH", "1") != "0"
    pf = DevicePrefetcher(dev)
    if prefetch:
        pf.feed(data)

    def one():
        cur = data
        if prefetch:
            cur = pf.take()
            pf.feed(data)
        _, _, total = step.train_step(cur)
        return float(total.detach())                       # D2H read of the step's loss (the trainer logs it)

    def timed(fn, n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            last = fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, last

    launches_eager = None
    for w_ in range(max(warmup, 3)):                       # step 1 learns the gradient set, step 2 builds the overlap hooks, step 3 captures the encoder graph
        n_before = _lib.launch_count()
        one()
        if w_ == 1:
            launches_eager = _lib.launch_count() - n_before    # every kernel of the step launched eagerly (before any graph exists)
    graphed = bool(step._graph_loop is not None and step._graph_loop.graph is not None)
    n0 = _lib.launch_count()
    ms, loss = timed(one, steps)
    launches = (_lib.launch_count() - n0) / steps
    # all-reduce alone (same buckets, nothing to overlap with): what the overlap hides
    ar_ms, ar_bytes, ar_calls = None, 0, 0
    if world > 1 and step._reducer is not None and step._reducer.buckets:
        bk = step._reducer.buckets
        ar_bytes = step._reducer.stats["bytes"]
        ar_calls = len(bk)

        def ar_only():
            works = [dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, async_op=True) for b in bk]
            for w_ in works:
                w_.wait()
        ar_ms, _ = timed(ar_only, 10)
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    # the same step with TF32 operands everywhere (the frozen E2VID encoder's bf16 tensor-core operands switched off)
    dtypes = {"frozen_e2vid_encoder": "bf16" if e2vid_model.CONVLSTM_BF16 else "tf32",
              "frozen_teacher": "bf16" if im.TEACHER_BF16 else "tf32", "trainable_modules": "tf32", "accumulate": "fp32"}
    ms_tf32 = None
    if e2vid_model.CONVLSTM_BF16 or im.TEACHER_BF16:
        saved = (e2vid_model.CONVLSTM_BF16, im.TEACHER_BF16)
        e2vid_model.CONVLSTM_BF16, im.TEACHER_BF16 = False, False
        if step._graph_loop is not None:
            step._graph_loop.reset()                       # the captured encoder loop holds the bf16 kernels
        try:
            for _ in range(3):                             # two eager calls + the capture
                one()
            ms_tf32, _ = timed(one, steps)
        finally:
            e2vid_model.CONVLSTM_BF16, im.TEACHER_BF16 = saved
            if step._graph_loop is not None:
                step._graph_loop.reset()

    # ... and with bf16 operands in the frozen teacher as well (opt-in, OESS_TEACHER_DTYPE=bf16: informational)
    ms_all_bf16 = None
    if e2vid_model.CONVLSTM_BF16 and not im.TEACHER_BF16:
        im.TEACHER_BF16 = True
        try:
            for _ in range(3):                             # the teacher's CUDA graph is re-captured (its state key changed)
                one()
            ms_all_bf16, _ = timed(one, steps)
        finally:
            im.TEACHER_BF16 = False

    base = None
    if baseline_steps > 0:
        if rank == 0:
            base = _literal_baseline(e2vid, back, teacher, ev, frame, pl, sp, S, NF, Hc, dev, baseline_steps,
                                     (e2vid_model, im, sn), step)
        if world > 1:
            dist.barrier()
    if rank != 0:
        return None
    out = {"metric": "end-to-end pretrain step (frame2voxel): event-frames/s = 20 x samples/s", "n_gpus": world,
           "batch_per_gpu": B, "events_per_frame": events, "steps": steps, "ms_per_step": ms,
           "samples_per_s": world * B / ms * 1e3, "event_frames_per_s": world * B * NF / ms * 1e3, "loss": loss,
           "operand_dtypes": dtypes, "ms_per_step_tf32_operands": ms_tf32,
           "ms_per_step_with_bf16_teacher_optin": ms_all_bf16,
           "own_kernel_launches_per_step": launches_eager if launches_eager is not None else launches,
           "own_kernel_launches_per_step_outside_cuda_graphs": launches, "peak_mem_gb": peak_gb, "h2d_bytes_per_step": h2d,
           "h2d_prefetch_on_side_stream": prefetch,
           "encoder_loop_cuda_graph": graphed, "d2h_bytes_per_step": 4,
           "allreduce": {"backend": "nccl" if world > 1 else None, "bytes_per_step": ar_bytes, "calls_per_step": ar_calls,
                         "ms_alone": ar_ms, "overlapped_with_backward": world > 1}}
    if base is not None:
        out["torch_cudnn_literal_baseline"] = base
        out["speedup_vs_literal"] = base["ms_per_step"] / ms
    return out


def _literal_baseline(e2vid, back, teacher, ev, frame, pl, sp, S, NF, Hc, dev, steps, mods, step_obj):
    """The same step as the reference writes it, on torch / cuDNN (none of this repository's kernels), same GPU."""
    from .. import voxel
    e2vid_model, im, sn = mods
    flags = (e2vid_model.USE_TENSOR_CORES, im.USE_TENSOR_CORES, sn.TRAIN_ON_TENSOR_CORES)
    e2vid_model.USE_TENSOR_CORES, im.USE_TENSOR_CORES, sn.TRAIN_ON_TENSOR_CORES = False, False, False
    try:
        # the reference's loader hands over the DENSE event tensor; build it once (untimed) and keep it in pinned host memory
        x, y, t, p = (a.to(dev) for a in (ev.x, ev.y, ev.t, ev.p))
        grids = voxel.dsec_events_to_voxel_grid(x, y, t, p, ev.rectify_map, 5, frame_offsets=ev.frame_offsets.to(dev), mode="ordered")
        B = grids.shape[0] // NF
        dense_host = grids.view(B, NF * 5, 480, 640)[:, :, :Hc, :].contiguous().cpu().pin_memory()
        del grids, x, y, t, p
        opt_v = torch.optim.AdamW([q for q in back.parameters() if q.requires_grad], lr=5e-4, fused=True)
        opt_f = torch.optim.AdamW([q for q in teacher.parameters() if q.requires_grad], lr=5e-4, fused=True)

        def one():
            opt_v.zero_grad(set_to_none=True)
            opt_f.zero_grad(set_to_none=True)
            back.train(); teacher.train()
            event = dense_host.to(dev, non_blocking=True)                                   # pretrain_trainer.py:428
            total, _, _ = literal_step(e2vid, back, teacher, event, frame.to(dev, non_blocking=True),
                                       pl.to(dev, non_blocking=True), sp.to(dev, non_blocking=True), S, NF)
            total.backward()
            opt_v.step(); opt_f.step()
            return float(total.detach())

        one()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "event_frames_per_s": B * NF / ms * 1e3, "loss": last,
                "steps": steps, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                "h2d_bytes_per_step": int(dense_host.numel() * 4 + frame.numel() * 4 + pl.numel() * 8 + sp.numel() * 8),
                "formulation": "pretrain_trainer.py:427-472 + loss_functions.py as written: dense event tensor from pinned host "
                               "memory, cuDNN (TF32 convolutions: torch default) / cuBLAS / torch.sparse, fused AdamW"}
    finally:
        e2vid_model.USE_TENSOR_CORES, im.USE_TENSOR_CORES, sn.TRAIN_ON_TENSOR_CORES = flags
