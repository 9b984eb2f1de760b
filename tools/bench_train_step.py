#!/usr/bin/env python
"""End-to-end OpenESS pretraining step (BASELINE config 3 per-GPU shard): CLI over
openess_b200/training/bench_step.py (the same measurement bench.py embeds as its `train_step` block).

    python tools/bench_train_step.py [--batch 4] [--steps 5] [--warmup 2] [--baseline-steps 3]     # torchrun for N > 1
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--events", type=int, default=100_000)
    ap.add_argument("--baseline-steps", type=int, default=0, help="also time the literal torch / cuDNN formulation on rank 0")
    args = ap.parse_args()
    import torch
    from openess_b200 import parallel
    from openess_b200.training import bench_step
    rank, local, world = parallel.init()
    res = bench_step.run(args.batch, args.steps, args.warmup, args.events, args.baseline_steps, rank, local, world)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
