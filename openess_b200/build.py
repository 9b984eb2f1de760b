"""Builds libopeness_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m openess_b200.build [--force]

The library is the product: there is no CPU fallback and no JIT.  The .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(CSRC, "_obj")
LIB = os.path.join(LIBDIR, "libopeness_b200.so")

# --fmad=false for the bit-exact voxelisers: every float op must round like the reference's separate
# numpy / torch ops (no FMA contraction).  The remaining files keep FMA.
SOURCES = {
    "common.cu": [],
    "voxel_trilinear.cu": ["--fmad=false"],
    "voxel_tbilinear.cu": ["--fmad=false"],
    "dsec_prestep.cu": ["--fmad=false"],
    "normalize.cu": ["--fmad=false"],
    "losses.cu": [],
    "infonce.cu": [],
    "pointwise.cu": [],
    "pixel_linear.cu": [],
    "tc_gemm.cu": [],
    "tc_convlstm.cu": [],
    "tc_conv.cu": [],
    "tc_wgrad.cu": [],
    "bn_nhwc.cu": [],
    "upnorm_pool.cu": [],
    "vit.cu": [],
    "tc_mha.cu": [],
    "pool_nhwc.cu": [],
    "resize.cu": [],
    "augment.cu": ["--fmad=false"],
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--threads", "0"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libopeness_b200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "openess_b200.h"))
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose or "warning" in out.lower():
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
