"""torch.profiler breakdown (top kernels by device time) of tools/bench_train_step.py -- used to find where the end-to-end
pretraining step spends its time (profiles/README.md)."""
import sys, os, json
sys.argv = ["bench_train_step.py", "--batch", "4", "--steps", "2", "--warmup", "2"]
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import bench_train_step as b
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    b.main()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:28]
tot = sum(e.device_time_total for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA)
for e in rows:
    print(f"{e.key[:90]:90s} {e.count:6d} {e.device_time_total/1e3:9.2f} ms")
