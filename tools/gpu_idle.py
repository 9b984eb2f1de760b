"""GPU idle time inside one pretraining step: union of the device-side kernel / memcpy intervals of a torch.profiler trace against
the step's wall span (are we launch-bound anywhere?)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200.training import bench_step  # noqa: E402

# warm everything up outside the profile, then profile a short run and analyse its LAST step
bench_step.run(batch=4, steps=1, warmup=2)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    res = bench_step.run(batch=4, steps=3, warmup=2)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
# the last 40 ms of device activity before the TF32 re-run would be ambiguous: take the busiest 45 ms window ending at each
# AdamW kernel instead: simply analyse gaps over the whole trace and report the distribution of idle gaps > 20 us
gaps, busy_end = [], iv[0][1]
for s, e, n in iv[1:]:
    if s > busy_end:
        gaps.append((s - busy_end, n))
    busy_end = max(busy_end, e)
total_span = iv[-1][1] - iv[0][0]
idle = sum(g for g, _ in gaps)
print(f"trace span {total_span / 1e3:.1f} ms, device idle {idle / 1e3:.1f} ms ({100 * idle / total_span:.1f} %), step {res['ms_per_step']:.2f} ms")
big = sorted(gaps, reverse=True)[:12]
for g, n in big:
    print(f"  gap {g:8.0f} us before {n[:90]}")
small = [g for g, _ in gaps if g < 50]
mid = [g for g, _ in gaps if 50 <= g < 2000]
print(f"gaps < 50 us: {len(small)} totalling {sum(small) / 1e3:.2f} ms; 50 us .. 2 ms: {len(mid)} totalling {sum(mid) / 1e3:.2f} ms")
