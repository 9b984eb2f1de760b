"""torch.profiler breakdown of the end-to-end pretraining step (openess_b200/training/bench_step.py): top kernels by device
time, and the Python call sites of the copy / conversion ops (`--shapes`: grouped by input shape) -- used to find where the step spends its time."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openess_b200.training import bench_step  # noqa: E402

stacks = "--shapes" in sys.argv
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=stacks) as prof:
    res = bench_step.run(batch=4, steps=2, warmup=2)
if stacks:
    rows = [e for e in prof.key_averages(group_by_input_shape=True)
            if e.key in ("aten::copy_", "aten::_to_copy", "aten::clone", "aten::contiguous", "aten::cat", "aten::upsample_nearest2d")]
    for e in sorted(rows, key=lambda e: -e.device_time_total)[:40]:
        print(f"{e.key:18s} {e.count:6d} {e.device_time_total / 1e3 / 4:8.3f} ms/step  {str(e.input_shapes)[:120]}")
else:
    # device kernels / memcpys only (self device time), per step; the build + warm-up share the profile, hence the divisor
    rows = sorted((e for e in prof.key_averages() if e.self_device_time_total > 0), key=lambda e: -e.self_device_time_total)
    total = sum(e.self_device_time_total for e in rows)
    print(f"device-busy time {total / 1e3 / 4:.2f} ms/step over {len(rows)} kernels")
    for e in rows[:40]:
        print(f"{e.key[:100]:100s} {e.count:6d} {e.self_device_time_total / 1e3 / 4:9.3f} ms/step")
print(res["ms_per_step"])
