// common.cuh -- shared device/host helpers for libopeness_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/openess_b200.h"

#define OESS_API extern "C" __attribute__((visibility("default")))

#define OESS_LAUNCH_CHECK()                          \
    do {                                             \
        cudaError_t e__ = cudaGetLastError();        \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

// Launch a kernel (or a memset) under the launch counter / optional event recorder, then check it.
#define OESS_KERNEL(name, st, ...)                   \
    do {                                             \
        {                                            \
            oess::prof::Scope ps__(name, st);        \
            __VA_ARGS__;                             \
        }                                            \
        OESS_LAUNCH_CHECK();                         \
    } while (0)

#define OESS_CUDA(call)                              \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

namespace oess {

constexpr int kNumSMs = 148;  // B200

// Launch bookkeeping: a process-wide launch counter (always on, one relaxed atomic per launch) and an
// optional per-host-thread recorder that brackets every kernel with CUDA events on its stream
// (oess_profile_begin / oess_profile_end).  Used by bench.py for `gpu_launches` and per-kernel times.
namespace prof {
struct Scope {
    const char* name;
    cudaStream_t st;
    void* rec;
    Scope(const char* name, cudaStream_t st);
    ~Scope();
};
}  // namespace prof

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace (host-side arithmetic only).
struct WsCarver {
    char* base;
    size_t off = 0;
    explicit WsCarver(void* p) : base((char*)p) {}
    template <class T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
    size_t total() const { return align_up(off, 256); }
};

// x86 cvttss2si / cvttsd2si semantics (what torch `.int()` and numpy `.astype(int64)` give on the
// reference's hosts): out-of-range and NaN become INT_MIN.  representations.py:27-29, data_util.py:74-81.
__device__ __forceinline__ int cvtt_f32_i32(float v) {
    return (v >= -2147483648.0f && v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
}
__device__ __forceinline__ long long cvtt_f64_i64(double v) {
    return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? __double2ll_rz(v)
                                                                      : (long long)0x8000000000000000LL;
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Streaming loads: event arrays are read once per pass.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

// ---------------------------------------------------------------------------------------------
// Chunk map: frame f owns chunks [chunk_start[f], chunk_start[f+1]) of `chunk` events each.
// One CTA; F is small (<= a few thousand frames per launch).
// ---------------------------------------------------------------------------------------------
__global__ void k_chunk_map(const int64_t* __restrict__ frame_offsets, int F, int chunk,
                            int* __restrict__ chunk_start);

// largest f in [0, F) with chunk_start[f] <= g, or -1 if g >= chunk_start[F]
__device__ __forceinline__ int find_frame(const int* __restrict__ chunk_start, int F, int g) {
    if (g >= chunk_start[F]) return -1;
    int lo = 0, hi = F;  // invariant: chunk_start[lo] <= g < chunk_start[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_start[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// Segmented exclusive scan of uint32 arrays, one CTA (1024 threads) per segment.
//   segment s = data[s * stride, s * stride + len)   (stride % 4 == 0)
__global__ void k_seg_exscan_u32(uint32_t* __restrict__ data, int64_t stride, int64_t len);

}  // namespace oess
