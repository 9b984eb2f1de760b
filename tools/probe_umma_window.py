#!/usr/bin/env python
"""Runs tools/probe_umma_window.cu on cuda:0 and prints, per (row0, stride byte offset, base offset), whether the tcgen05 product
over a shifted window of a TMA-written 128-byte-swizzled tile equals the expected one (small integers: exact in TF32)."""
import ctypes
import json
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_probe", "libprobe.so")


def build():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           "-o", LIB, os.path.join(HERE, "probe_umma_window.cu"), "-lcudart"])


if __name__ == "__main__":
    if not os.path.exists(LIB):
        build()
    lib = ctypes.CDLL(LIB)
    lib.probe_umma_window.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randint(-8, 9, (180, 32), device="cuda", generator=g).float()
    B = torch.randint(-8, 9, (64, 32), device="cuda", generator=g).float()
    for row0, sbo, boff in [(0, 1024, 0), (0, 1280, 0), (8, 1280, 0), (11, 1280, 0), (11, 1280, 3), (3, 1024, 0), (3, 1024, 3), (21, 1280, 0)]:
        D = torch.zeros(128, 64, device="cuda")
        rc = lib.probe_umma_window(A.data_ptr(), B.data_ptr(), D.data_ptr(), row0, sbo, boff)
        m = torch.arange(128, device="cuda")
        rows = row0 + (m // 8) * (sbo // 128) + m % 8
        ref = A[rows] @ B.t()
        print(json.dumps({"row0": row0, "sbo_bytes": sbo, "base_offset": boff, "rc": rc, "exact": bool(torch.equal(D, ref)),
                          "rows_wrong": int((D != ref).any(1).sum())}))
