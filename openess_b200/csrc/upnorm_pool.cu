// upnorm_pool.cu -- the tail of the frame-branch teacher fused with the superpixel pooling that consumes it:
//     q_sum[m] = sum_{pixels of superpixel m}  normalize_C( bilinear_x4_align_corners( d ) )[pixel]
// Replaces models/image_model.py:121-124,139-141 (nn.Upsample(scale_factor=4, bilinear, align_corners=True) +
// F.normalize(p=2, dim=1)) followed by training/pretrain_trainer.py:446-463 (sparse one-hot matmul on a permuted copy).
// The reference materialises the [B, 256, H, W] map (288 MB / sample at 440 x 640) and passes over it ~8 times forward
// and backward; here it never exists: the kernels read the LOW-resolution decoder output d [B, h, w, C] (channels-last,
// 18 MB / sample) and the superpixel ids, and write [M, C] sums (forward) / the [B, h, w, C] gradient (backward).
//
// One warp walks a run of 64 consecutive output pixels of one row; lane l owns channels [8 l, 8 l + 8) (C = 256).
// The four bilinear neighbours stay in registers while the source column is unchanged (4 output pixels per source
// column), the per-pixel L2 norm is one warp reduction, and sums for one superpixel are kept in registers until the id
// changes (superpixels are spatially coherent), then flushed with red.global.add.f32.  The backward pass recomputes the
// normalised vector, applies the normalise Jacobian, and accumulates the four neighbour gradients in registers, flushing a
// source column once the run has passed it.  HBM-bound: ~8 B / output pixel (ids) + the low-resolution map.
#include "common.cuh"

namespace oess {

constexpr int kUpC = 256;        // channels (model_n_out, image_model.py:96)
constexpr int kUpCPL = 8;        // channels per lane
constexpr int kUpRun = 64;       // output pixels per warp task
constexpr int kUpWarps = 4;

struct Vec8 {
    float v[8];
};
__device__ __forceinline__ Vec8 ld8(const float* p) {
    Vec8 r;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void red8(float* p, const Vec8& a) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(p + j, a.v[j]);
}

// torch upsample_bilinear2d, align_corners = True (ATen UpSample.cuh area_pixel_compute_source_index): src = scale * dst
struct Src { int i0, i1; float l0, l1; };
__device__ __forceinline__ Src src_index(int dst, float scale, int in_size) {
    const float s = scale * (float)dst;
    Src r;
    r.i0 = (int)s;
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = s - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

template <bool BWD>
__global__ void __launch_bounds__(kUpWarps * 32)
k_upnorm_pool(const float* __restrict__ d, const int64_t* __restrict__ seg, const float* __restrict__ g_sum, int B, int h,
              int w, int H, int W, int S, int64_t M, float sy, float sx, float* __restrict__ pooled,
              float* __restrict__ counts, float* __restrict__ d_grad, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int runs = (W + kUpRun - 1) / kUpRun;
    const int64_t task = (int64_t)blockIdx.x * kUpWarps + warp;
    if (task >= (int64_t)B * H * runs) return;
    const int b = (int)(task / ((int64_t)H * runs));
    const int rem = (int)(task - (int64_t)b * H * runs);
    const int y = rem / runs, x_lo = (rem - y * runs) * kUpRun, x_hi = min(x_lo + kUpRun, W);
    const Src ys = src_index(y, sy, h);
    const float* row0 = d + (((int64_t)b * h + ys.i0) * w) * kUpC + lane * kUpCPL;
    const float* row1 = d + (((int64_t)b * h + ys.i1) * w) * kUpC + lane * kUpCPL;
    const int64_t* srow = seg + ((int64_t)b * H + y) * W;

    int cx0 = -1, cx1 = -1;                         // source columns held in registers
    Vec8 tl, tr, bl, br;                            // neighbours: (row i0 / i1) x (col cx0 / cx1)
    Vec8 acc;                                       // fwd: running superpixel sum
    Vec8 gtl, gtr, gbl, gbr;                        // bwd: neighbour gradients of the current column pair
    Vec8 gs;                                        // bwd: g_sum row of the current superpixel
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.v[j] = 0.f; gtl.v[j] = gtr.v[j] = gbl.v[j] = gbr.v[j] = 0.f; gs.v[j] = 0.f; tl.v[j] = tr.v[j] = bl.v[j] = br.v[j] = 0.f; }
    int64_t cur = -1;
    float cnt = 0.f;

    for (int x = x_lo; x < x_hi; ++x) {
        const Src xs = src_index(x, sx, w);
        if (xs.i0 != cx0) {
            if (BWD && cx0 >= 0) {                  // the run has passed source column cx0: flush its gradients
                float* g0 = d_grad + (((int64_t)b * h + ys.i0) * w + cx0) * kUpC + lane * kUpCPL;
                float* g1 = d_grad + (((int64_t)b * h + ys.i1) * w + cx0) * kUpC + lane * kUpCPL;
                red8(g0, gtl);
                red8(g1, gbl);
                if (cx1 != cx0 && xs.i0 == cx1) {   // shift right column to the left
                    gtl = gtr; gbl = gbr;
                } else {                            // jump (cannot happen for scale >= 1, kept for safety)
                    if (cx1 != cx0) { red8(g0 + (int64_t)(cx1 - cx0) * kUpC, gtr); red8(g1 + (int64_t)(cx1 - cx0) * kUpC, gbr); }
#pragma unroll
                    for (int j = 0; j < 8; ++j) gtl.v[j] = gbl.v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) gtr.v[j] = gbr.v[j] = 0.f;
            }
            if (xs.i0 == cx1 && cx1 != cx0) { tl = tr; bl = br; }
            else { tl = ld8(row0 + (int64_t)xs.i0 * kUpC); bl = ld8(row1 + (int64_t)xs.i0 * kUpC); }
            cx0 = xs.i0;
            cx1 = -2;                               // force the right column to be (re)loaded below
        }
        if (xs.i1 != cx1) {
            if (xs.i1 == cx0) { tr = tl; br = bl; }
            else { tr = ld8(row0 + (int64_t)xs.i1 * kUpC); br = ld8(row1 + (int64_t)xs.i1 * kUpC); }
            cx1 = xs.i1;
        }
        // bilinear value (UpSample: h0 (w0 TL + w1 TR) + h1 (w0 BL + w1 BR)) and L2 normalisation over channels
        Vec8 v;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v.v[j] = ys.l0 * (xs.l0 * tl.v[j] + xs.l1 * tr.v[j]) + ys.l1 * (xs.l0 * bl.v[j] + xs.l1 * br.v[j]);
            ss += v.v[j] * v.v[j];
        }
        const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);          // F.normalize eps
        const float inv = 1.0f / nrm;
        int64_t id = __ldg(srow + x) + (int64_t)b * S;                 // pretrain_trainer.py:446-449
        if (id < 0 || id >= M) { if (status) *status = 1; id = -1; }
        if (!BWD) {
            if (id != cur) {
                if (cur >= 0) { red8(pooled + cur * kUpC + lane * kUpCPL, acc); if (lane == 0) atomicAdd(counts + cur, cnt); }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
                cnt = 0.f;
                cur = id;
            }
            if (id >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j] * inv;
                cnt += 1.0f;
            }
        } else {
            if (id != cur) {
                cur = id;
                if (id >= 0) gs = ld8(g_sum + id * kUpC + lane * kUpCPL);
            }
            if (id >= 0) {
                // u = v / nrm;  dL/dv = (g - u (u . g)) / nrm      (normalise Jacobian; clamp branch: g / eps)
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) dot += v.v[j] * inv * gs.v[j];
                dot = warp_sum(dot);
                const bool clamped = nrm <= 1e-12f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gv = clamped ? gs.v[j] * inv : (gs.v[j] - v.v[j] * inv * dot) * inv;
                    gtl.v[j] += gv * ys.l0 * xs.l0;
                    gbl.v[j] += gv * ys.l1 * xs.l0;
                    if (cx1 == cx0) { gtl.v[j] += gv * ys.l0 * xs.l1; gbl.v[j] += gv * ys.l1 * xs.l1; }
                    else { gtr.v[j] += gv * ys.l0 * xs.l1; gbr.v[j] += gv * ys.l1 * xs.l1; }
                }
            }
        }
    }
    if (!BWD) {
        if (cur >= 0) { red8(pooled + cur * kUpC + lane * kUpCPL, acc); if (lane == 0) atomicAdd(counts + cur, cnt); }
    } else if (cx0 >= 0) {
        float* g0 = d_grad + (((int64_t)b * h + ys.i0) * w + cx0) * kUpC + lane * kUpCPL;
        float* g1 = d_grad + (((int64_t)b * h + ys.i1) * w + cx0) * kUpC + lane * kUpCPL;
        red8(g0, gtl);
        red8(g1, gbl);
        if (cx1 != cx0) { red8(g0 + (int64_t)(cx1 - cx0) * kUpC, gtr); red8(g1 + (int64_t)(cx1 - cx0) * kUpC, gbr); }
    }
}

}  // namespace oess

using namespace oess;

static int upnorm_check(const void* d, const void* seg, int B, int h, int w, int C, int H, int W, int64_t M) {
    if (!d || !seg || B <= 0 || h <= 0 || w <= 0 || H < h || W < w || M <= 0) return OESS_E_ARG;
    if (C != kUpC) return OESS_E_ARG;
    if ((uintptr_t)d & 15) return OESS_E_ARG;
    if ((int64_t)B * H * ((W + kUpRun - 1) / kUpRun) / kUpWarps + 1 >= (1ll << 31)) return OESS_E_RANGE;
    return OESS_OK;
}
static float up_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f; }

// d: [B, h, w, 256] channels-last; seg: int64 [B, H, W] superpixel ids (b * S is added inside); pooled_sum [M, 256] and
// counts [M] are ZEROED here and receive sums / pixel counts; status (optional): set to 1 if an id falls outside [0, M).
OESS_API int oess_upnorm_pool_fwd(const float* d, const int64_t* seg, int B, int h, int w, int C, int H, int W, int S,
                                  int64_t M, float* pooled_sum, float* counts, int32_t* status, oess_stream_t stream) {
    int rc = upnorm_check(d, seg, B, h, w, C, H, W, M);
    if (rc) return rc;
    if (!pooled_sum || !counts) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(pooled_sum, 0, sizeof(float) * (size_t)M * kUpC, st));
    OESS_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * (size_t)M, st));
    const int64_t tasks = (int64_t)B * H * ((W + kUpRun - 1) / kUpRun);
    OESS_KERNEL("upnorm_pool_fwd", st, k_upnorm_pool<false><<<(unsigned)((tasks + kUpWarps - 1) / kUpWarps), kUpWarps * 32, 0, st>>>(
        d, seg, nullptr, B, h, w, H, W, S, M, up_scale(h, H), up_scale(w, W), pooled_sum, counts, nullptr, status));
    return OESS_OK;
}

// g_sum: [M, 256] gradient w.r.t. pooled_sum; d_grad: [B, h, w, 256] channels-last, ZEROED here.
OESS_API int oess_upnorm_pool_bwd(const float* d, const int64_t* seg, const float* g_sum, int B, int h, int w, int C, int H,
                                  int W, int S, int64_t M, float* d_grad, oess_stream_t stream) {
    int rc = upnorm_check(d, seg, B, h, w, C, H, W, M);
    if (rc) return rc;
    if (!g_sum || !d_grad || ((uintptr_t)g_sum & 15)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(float) * (size_t)B * h * w * kUpC, st));
    const int64_t tasks = (int64_t)B * H * ((W + kUpRun - 1) / kUpRun);
    OESS_KERNEL("upnorm_pool_bwd", st, k_upnorm_pool<true><<<(unsigned)((tasks + kUpWarps - 1) / kUpWarps), kUpWarps * 32, 0, st>>>(
        d, seg, g_sum, B, h, w, H, W, S, M, up_scale(h, H), up_scale(w, W), nullptr, nullptr, d_grad, nullptr));
    return OESS_OK;
}
