// radix.cuh -- batched, STABLE least-significant-digit radix sort of per-frame event records.
//
// Why it exists: the reference accumulates every voxel sequentially in event order (np.add.at /
// serial put_).  To replay that order on a GPU the events of each frame are
// stably sorted by the pixel they fall on; a gather kernel then walks each pixel's (short) event
// list in original order.  Stability is the whole point, so ranking inside a CTA is done in item
// order (warp-synchronous match/popc ranking over consecutive items, then a warp-major prefix).
//
// Geometry: 10-bit digits (1024 bins), 256 threads x 8 items = 2048-event chunks.  A frame with n
// events owns ceil(n / 2048) chunks; chunk_start[] (k_chunk_map) maps a flat CTA index to
// (frame, chunk).  One pass = k_hist -> k_prefix_chunks -> k_bin_scan -> k_scatter.
#pragma once
#include <stdlib.h>
#include "common.cuh"

namespace oess {
namespace radix {

constexpr int kBits = 10;
constexpr int kBins = 1 << kBits;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItemsPerThread = 8;
constexpr int kChunk = kThreads * kItemsPerThread;  // 2048

static inline int64_t max_chunks(int64_t n_total, int F) { return (n_total + kChunk - 1) / kChunk + F; }

// Src concept:
//   typedef ... Item;
//   __device__ Item     load(int f, int64_t fbeg, uint32_t li) const;   // li = index inside the frame
//   __device__ uint32_t key(const Item&) const;                          // full sort key
//   __device__ uint32_t key_at(int f, int64_t fbeg, uint32_t li) const;  // == key(load(f, fbeg, li))
//
// hist: [total_chunks, kBins] chunk-major.  pix (optional): per-frame full-key counts for the CSR
// offsets used by the gather kernels (pix[f * pix_stride + key] += 1).
template <class Src>
__global__ void __launch_bounds__(kThreads)
k_hist(Src src, const int64_t* __restrict__ frame_offsets, const int* __restrict__ chunk_start, int F,
       int shift, uint32_t mask, uint32_t* __restrict__ hist, uint32_t* __restrict__ pix,
       int64_t pix_stride, int nb) {
    __shared__ uint32_t s_hist[kBins];
    const int g = blockIdx.x;
    const int f = find_frame(chunk_start, F, g);
    if (f < 0) return;
    const int c = g - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const uint32_t nf = (uint32_t)(frame_offsets[f + 1] - fbeg);
    for (int b = threadIdx.x; b < nb; b += kThreads) s_hist[b] = 0;      // nb = bins that can occur (multiple of 128)
    __syncthreads();
    const uint32_t cbeg = (uint32_t)c * kChunk;
#pragma unroll
    for (int s = 0; s < kItemsPerThread; ++s) {
        const uint32_t li = cbeg + s * kThreads + threadIdx.x;
        if (li < nf) {
            const uint32_t key = src.key_at(f, fbeg, li);   // may read less than a full record
            atomicAdd(&s_hist[(key >> shift) & mask], 1u);
            if (pix) atomicAdd(&pix[(int64_t)f * pix_stride + key], 1u);
        }
    }
    __syncthreads();
    uint32_t* h = hist + (int64_t)g * kBins;
    for (int b = threadIdx.x; b < nb; b += kThreads) h[b] = s_hist[b];
}

// Exclusive prefix over the chunks of each frame, per bin (in place); per-frame bin totals to tot.
// grid (F, kBins / 128), 128 threads.
__global__ void k_prefix_chunks(uint32_t* __restrict__ hist, const int* __restrict__ chunk_start,
                                uint32_t* __restrict__ tot, int nb);
// Exclusive scan over the kBins totals of each frame (in place).  grid F, kBins threads.
__global__ void k_bin_scan(uint32_t* __restrict__ tot);

template <class Src>
__global__ void __launch_bounds__(kThreads, 4)
k_scatter(Src src, const int64_t* __restrict__ frame_offsets, const int* __restrict__ chunk_start, int F,
          int shift, uint32_t mask, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ binbase,
          typename Src::Item* __restrict__ dst, int nb) {
    __shared__ uint32_t s_cnt[kWarps][kBins];
    const int g = blockIdx.x;
    const int f = find_frame(chunk_start, F, g);
    if (f < 0) return;
    const int c = g - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const uint32_t nf = (uint32_t)(frame_offsets[f + 1] - fbeg);
    for (int b = threadIdx.x; b < nb; b += kThreads) {
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) s_cnt[ww][b] = 0;
    }
    __syncthreads();

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const uint32_t wbeg = (uint32_t)c * kChunk + (uint32_t)w * (32 * kItemsPerThread);
    typename Src::Item it[kItemsPerThread];
    uint32_t dig[kItemsPerThread];
    uint32_t rk[kItemsPerThread];
    // Warp w owns items [wbeg, wbeg + 256) of the frame; step s covers 32 consecutive items, so
    // (step, lane) order == event order and the ranking below is stable.
#pragma unroll
    for (int s = 0; s < kItemsPerThread; ++s) {
        const uint32_t li = wbeg + s * 32 + lane;
        const bool act = li < nf;
        uint32_t d = kBins;  // sentinel digit for the ragged tail
        if (act) {
            it[s] = src.load(f, fbeg, li);
            d = (src.key(it[s]) >> shift) & mask;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && act) {
            old = s_cnt[w][d];
            s_cnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        dig[s] = d;
        rk[s] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // warp-major exclusive prefix per bin, seeded with the global base of (frame, bin, chunk)
    const uint32_t* h = hist + (int64_t)g * kBins;
    const uint32_t* bb = binbase + (int64_t)f * kBins;
    for (int b = threadIdx.x; b < nb; b += kThreads) {
        uint32_t run = bb[b] + h[b];
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            const uint32_t v = s_cnt[ww][b];
            s_cnt[ww][b] = run;
            run += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kItemsPerThread; ++s) {
        if (dig[s] < (uint32_t)kBins) dst[fbeg + s_cnt[w][dig[s]] + rk[s]] = it[s];
    }
}

// Host-side driver for one pass.  `tot` is [F, kBins]; `hist` is [max_chunks, kBins].
template <class Src>
static inline int run_pass(const Src& src, const int64_t* frame_offsets, const int* chunk_start, int F,
                           int64_t n_chunks_ub, int shift, uint32_t mask, uint32_t* hist, uint32_t* tot,
                           uint32_t* pix, int64_t pix_stride, typename Src::Item* dst, cudaStream_t st, int n_keys = kBins) {
    // bins that can occur in this pass, rounded up to the 128-bin granularity of k_prefix_chunks: the per-CTA zeroing /
    // prefix loops and the chunk-histogram traffic shrink with it (482 of 1024 bins for the 480-row DSEC sensor)
    int nb = (int)(((int64_t)(n_keys < (int)(mask + 1) ? n_keys : (int)(mask + 1)) + 127) / 128 * 128);
    if (nb > kBins) nb = kBins;
    if (n_chunks_ub <= 0 || F <= 0) return 0;
    OESS_KERNEL("k_hist", st, k_hist<Src><<<(unsigned)n_chunks_ub, kThreads, 0, st>>>(src, frame_offsets, chunk_start, F, shift, mask,
                                                          hist, pix, pix_stride, nb));
    OESS_KERNEL("k_prefix_chunks", st, k_prefix_chunks<<<dim3((unsigned)F, kBins / 128), 128, 0, st>>>(hist, chunk_start, tot, nb));
    OESS_KERNEL("k_bin_scan", st, k_bin_scan<<<(unsigned)F, kBins, 0, st>>>(tot));
    OESS_KERNEL("k_scatter", st, k_scatter<Src><<<(unsigned)n_chunks_ub, kThreads, 0, st>>>(src, frame_offsets, chunk_start, F,
                                                             shift, mask, hist, tot, dst, nb));
    return 0;
}

static inline int key_bits(uint32_t n_keys) {  // bits needed to represent keys 0 .. n_keys-1
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long)n_keys) ++b;
    return b;
}

}  // namespace radix
}  // namespace oess
