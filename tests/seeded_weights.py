"""Deterministic, platform-independent weights for model goldens that are too large to commit (numpy Generator streams
are stable across numpy versions and platforms, torch's are not guaranteed to be).  Used by oracle/make_golden_models.py
(to load the REFERENCE module) and by the tests (to load the mirror): both get bit-identical state dicts."""
import numpy as np
import torch


def seeded_state_dict(module, seed):
    rng = np.random.default_rng(seed)
    sd = {}
    for name, t in module.state_dict().items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros_like(t)
            continue
        if name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            v = rng.normal(0, 0.1, shape)
        elif t.ndim == 1 and name.endswith("weight"):          # norm-layer gains (conv / linear weights are >= 2-D)
            v = rng.uniform(0.8, 1.2, shape)
        elif t.ndim == 1:
            v = rng.normal(0, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rng.normal(0, 1.0 / np.sqrt(fan_in), shape)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return sd
