// dsec_prestep.cu -- raw DSEC records -> the four float32 arrays VoxelGrid.convert consumes.
// Fuses DSEC/dataset/sequence_ov.py:204-210 (rectify_events: gather rectify_map[y, x]) with
// sequence_ov.py:154-159 (events_to_voxel_grid: t = f32(t - t[0]); t = t / t[-1]; pol = f32(p)),
// per frame of a batch.  HBM-bound streaming kernel: 9 B (u16 x, u16 y, u32 t, u8 p; the on-disk DSEC
// layout, DSEC/utils/eventslicer.py) or 13 B (int64 t) in + 16 B out per event; the 2.46 MB rectify map stays
// L2-resident (8 B gather per event).
// Compile with --fmad=false (the float division must round like numpy's).
#include "common.cuh"

namespace oess {

// The reference adds the file's t_offset to the uint32 timestamps (int64 microseconds) and then only uses
// differences inside a frame, so uint32 input gives the same result as long as a frame does not wrap.
template <class TT>
__global__ void __launch_bounds__(256)
k_dsec_rectify_tnorm(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const TT* __restrict__ t,
                     const uint8_t* __restrict__ p, const float2* __restrict__ rectify_map,
                     const int64_t* __restrict__ frame_offsets, int64_t n, int F, int H, int W,
                     float* __restrict__ xo, float* __restrict__ yo, float* __restrict__ po,
                     float* __restrict__ to, int32_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // frame of event i: frame_offsets[lo] <= i < frame_offsets[lo + 1].  Almost every block of 256 consecutive events lies inside
    // one frame (100 000 events per frame): two threads search for the block's first and last event, the others search only the
    // (usually empty) range between the two answers instead of walking log2(F) dependent loads each.
    __shared__ int s_f[2];
    if (threadIdx.x < 2) {
        int64_t q = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x ? blockDim.x - 1 : 0);
        if (q > n - 1) q = n - 1;
        int lo = 0, hi = F;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (frame_offsets[mid] <= q) lo = mid; else hi = mid;
        }
        s_f[threadIdx.x] = lo;
    }
    __syncthreads();
    if (i >= n) return;
    int lo = s_f[0], hi = s_f[1] + 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (frame_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t fbeg = frame_offsets[lo], fend = frame_offsets[lo + 1];
    const long long t0 = (long long)t[fbeg];
    // int64 differences are < 2^53, so int64 -> f32 equals the reference's f64 -> f32 cast (:155)
    const float tlast = __ll2float_rn((long long)t[fend - 1] - t0);
    const float ti = __ll2float_rn((long long)__ldcs(t + i) - t0);
    const unsigned xi = x[i], yi = y[i];
    float2 m = make_float2(0.f, 0.f);
    if (xi < (unsigned)W && yi < (unsigned)H) {
        m = __ldg(rectify_map + (int64_t)yi * W + xi);      // :210 rectify_map[y, x]
    } else if (status) {
        *status = 1;                                        // :208-209 asserts
    }
    __stcs(xo + i, m.x);
    __stcs(yo + i, m.y);
    __stcs(po + i, (float)p[i]);                            // :159
    __stcs(to + i, __fdiv_rn(ti, tlast));                   // :156 (no guard: 0/0 -> NaN, like the reference)
}

template <class TT>
static int run_prestep(const uint16_t* x, const uint16_t* y, const TT* t, const uint8_t* p, const float* rectify_map,
                       const int64_t* frame_offsets, int64_t n, int F, int H, int W, float* xo, float* yo, float* po,
                       float* to, int32_t* status, cudaStream_t st) {
    if (n < 0 || F < 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    if (n == 0 || F == 0) return OESS_OK;
    if (!x || !y || !t || !p || !rectify_map || !frame_offsets || !xo || !yo || !po || !to) return OESS_E_ARG;
    if (status) OESS_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    OESS_KERNEL("dsec_rectify_tnorm", st, k_dsec_rectify_tnorm<TT><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        x, y, t, p, reinterpret_cast<const float2*>(rectify_map), frame_offsets, n, F, H, W, xo, yo, po, to, status));
    return OESS_OK;
}

}  // namespace oess

using namespace oess;

OESS_API int oess_dsec_rectify_tnorm(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p,
                                     const float* rectify_map, const int64_t* frame_offsets, int64_t n, int F,
                                     int H, int W, float* xo, float* yo, float* po, float* to, int32_t* status,
                                     oess_stream_t stream) {
    return run_prestep<int64_t>(x, y, t, p, rectify_map, frame_offsets, n, F, H, W, xo, yo, po, to, status,
                                (cudaStream_t)stream);
}

OESS_API int oess_dsec_rectify_tnorm_u32(const uint16_t* x, const uint16_t* y, const uint32_t* t, const uint8_t* p,
                                         const float* rectify_map, const int64_t* frame_offsets, int64_t n, int F,
                                         int H, int W, float* xo, float* yo, float* po, float* to, int32_t* status,
                                         oess_stream_t stream) {
    return run_prestep<uint32_t>(x, y, t, p, rectify_map, frame_offsets, n, F, H, W, xo, yo, po, to, status,
                                 (cudaStream_t)stream);
}
