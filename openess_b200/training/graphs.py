"""CUDA-graph capture of the frozen recurrent encoder loop of the pretraining step.

`task_train_step` (training/pretrain_trainer.py:436-442) feeds the 20 event frames of a sample through E2VID one after the other
and keeps only the latents of the last one.  At batch 4 that is 20 x ~15 launches of small kernels (normalisation, head conv,
three strided convs, three ConvLSTM steps) issued from Python: the host needs longer to issue them than the GPU to run them
(`tools/gpu_idle.py`: ~5 ms of device idle time per 40 ms step).  The loop is a pure function of the event tensor with static
shapes, no autograd and no host synchronisation, so it is captured once (after two eager calls that build every cache) and
replayed; the event tensor lives in a persistent buffer the voxeliser writes into, the latents in the graph's private pool."""
import torch


class GraphedEncoderLoop:
    def __init__(self, reconstructor, nsteps, channels, eager_calls=2):
        self.rec, self.nsteps, self.C = reconstructor, int(nsteps), int(channels)
        self.eager_calls = int(eager_calls)
        self.graph, self.out, self.key, self.calls, self.disabled = None, None, None, 0, False

    def reset(self):
        """Forget the captured graph (the encoder's kernels / dtype switches changed): two eager calls, then a new capture."""
        self.graph, self.out, self.key, self.calls = None, None, None, 0

    def _loop(self, event):
        self.rec.last_states_for_each_channel = {"grayscale": None}
        latent = None
        for i in range(self.nsteps):
            _, _, latent = self.rec.update_reconstruction(event[:, i * self.C:(i + 1) * self.C])
        return latent

    def __call__(self, event):
        key = (event.data_ptr(), tuple(event.shape), tuple(event.stride()), event.dtype)
        if self.disabled or not event.is_cuda:
            return self._loop(event)
        if self.graph is not None and key == self.key:
            self.graph.replay()
            return self.out
        if key != self.key:                                    # new buffer / shape: start over with eager calls
            self.graph, self.out, self.key, self.calls = None, None, key, 0
        self.calls += 1
        if self.calls <= self.eager_calls:
            return self._loop(event)
        try:
            torch.cuda.synchronize(event.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._loop(event)
            self.graph, self.out = g, out
            g.replay()
            return out
        except Exception as exc:                               # capture not possible here: stay eager, say so once
            import warnings
            warnings.warn(f"GraphedEncoderLoop: CUDA-graph capture failed ({exc}); running the loop eagerly")
            self.disabled, self.graph, self.out = True, None, None
            torch.cuda.synchronize(event.device)
            return self._loop(event)


class GraphedCall:
    """`fn(x)` (a frozen, no-grad, static-shape CUDA computation returning a tensor) as a CUDA graph: two eager calls, capture on
    the third, then `static_in.copy_(x)` + replay.  `state()` (optional) returns anything whose change must drop the graph (module
    mode flags, parameter versions); `reset()` drops it explicitly."""

    def __init__(self, fn, state=None, eager_calls=2):
        self.fn, self.state, self.eager_calls = fn, state, int(eager_calls)
        self.disabled = False
        self.reset()

    def reset(self):
        self.graph, self.out, self.key, self.calls, self.static_in = None, None, None, 0, None

    def __call__(self, x):
        if self.disabled or not x.is_cuda or torch.cuda.is_current_stream_capturing():
            return self.fn(x)
        key = (tuple(x.shape), x.dtype, x.device, x.is_contiguous(memory_format=torch.channels_last),
               self.state() if self.state is not None else None)
        if self.graph is not None and key == self.key:
            self.static_in.copy_(x)
            self.graph.replay()
            return self.out
        if key != self.key:
            self.reset()
            self.key = key
        self.calls += 1
        if self.calls <= self.eager_calls:
            return self.fn(x)
        try:
            torch.cuda.synchronize(x.device)
            self.static_in = x.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.fn(self.static_in)
            self.graph, self.out = g, out
            g.replay()
            return out
        except Exception as exc:
            import warnings
            warnings.warn(f"GraphedCall: CUDA-graph capture failed ({exc}); running eagerly")
            self.disabled = True
            self.reset()
            torch.cuda.synchronize(x.device)
            return self.fn(x)
