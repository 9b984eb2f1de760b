// resize.cu -- bilinear resize (align_corners = False) of plane tensors, forward and backward:
//     F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False)        x [B, C, h, w] -> [B, C, H, W]
// the full-resolution upsampling of the DeepLabv3 logits and 256-channel features (models/deeplabv3.py:53-56 of the
// reference: x32 / x16, a 1.15 GB feature tensor at B = 4).  torch's backward scatters every output gradient into its four
// source pixels with atomics (37 ms per DeepLabv3 head step at B = 4); here the backward is a GATHER and separable:
//     tmp[b, c, y, j] = sum_x wx(x -> j) g[b, c, y, x]      (one pass over g, contiguous runs of ~2 W / w elements per thread)
//     dx [b, c, i, j] = sum_y wy(y -> i) tmp[b, c, y, j]
// with the forward's own weights (ATen UpSample.cuh area_pixel_compute_source_index: src = scale * (dst + 0.5) - 0.5 clamped
// at 0, i0 = floor(src), i1 = i0 + (i0 < in - 1), lambda1 = src - i0).  HBM-bound: 4 B / output element either way.
#include "common.cuh"

namespace oess {
namespace resize {

struct Tap { int i0, i1; float l0, l1; };

__device__ __forceinline__ Tap tap_of(int dst, float scale, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Tap t;
    t.i0 = min((int)src, in_size - 1);
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}

__global__ void __launch_bounds__(256)
k_bilinear_fwd(const float* __restrict__ x, int64_t planes, int h, int w, int H, int W, float sh, float sw, float* __restrict__ out) {
    const int64_t total = planes * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % W);
        const int yo = (int)((i / W) % H);
        const int64_t p = i / ((int64_t)W * H);
        const Tap ty = tap_of(yo, sh, h), tx = tap_of(xo, sw, w);
        const float* s = x + p * h * w;
        const float v00 = __ldg(s + ty.i0 * w + tx.i0), v01 = __ldg(s + ty.i0 * w + tx.i1);
        const float v10 = __ldg(s + ty.i1 * w + tx.i0), v11 = __ldg(s + ty.i1 * w + tx.i1);
        __stcs(out + i, ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11));
    }
}

// source index range [lo, hi] of destinations that can touch source `j`
__device__ __forceinline__ void dst_range(int j, float scale, int out_size, int& lo, int& hi) {
    const float inv = 1.0f / scale;
    lo = max(0, (int)floorf(((float)j - 1.0f + 0.5f) * inv - 0.5f) - 1);
    hi = min(out_size - 1, (int)ceilf(((float)j + 1.0f + 0.5f) * inv - 0.5f) + 1);
}

// pass 1: tmp [planes, H, w]
__global__ void __launch_bounds__(256)
k_bilinear_bwd_x(const float* __restrict__ g, int64_t rows, int w, int W, float sw, float* __restrict__ tmp) {
    const int64_t total = rows * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % w);
        const int64_t r = i / w;
        int lo, hi;
        dst_range(j, sw, W, lo, hi);
        const float* gr = g + r * W;
        float acc = 0.f;
        for (int xo = lo; xo <= hi; ++xo) {
            const Tap t = tap_of(xo, sw, w);
            const float v = __ldg(gr + xo);
            if (t.i0 == j) acc += t.l0 * v;
            if (t.i1 == j) acc += t.l1 * v;
        }
        tmp[i] = acc;
    }
}

// pass 2: dx [planes, h, w]
__global__ void __launch_bounds__(256)
k_bilinear_bwd_y(const float* __restrict__ tmp, int64_t planes, int h, int w, int H, float sh, float* __restrict__ dx) {
    const int64_t total = planes * h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % w);
        const int ii = (int)((i / w) % h);
        const int64_t p = i / ((int64_t)w * h);
        int lo, hi;
        dst_range(ii, sh, H, lo, hi);
        const float* tp = tmp + p * H * w + j;
        float acc = 0.f;
        for (int yo = lo; yo <= hi; ++yo) {
            const Tap t = tap_of(yo, sh, h);
            const float v = __ldg(tp + (int64_t)yo * w);
            if (t.i0 == ii) acc += t.l0 * v;
            if (t.i1 == ii) acc += t.l1 * v;
        }
        dx[i] = acc;
    }
}

static unsigned grid_for(int64_t total) {
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)kNumSMs * 32) g = (int64_t)kNumSMs * 32;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace resize
}  // namespace oess

using namespace oess;

// x [planes, h, w] -> out [planes, H, W] (planes = B * C of a contiguous NCHW tensor)
OESS_API int oess_bilinear_resize_planes(const float* x, int64_t planes, int h, int w, int H, int W, float* out,
                                         oess_stream_t stream) {
    if (!x || !out || planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("bilinear_resize_planes", st, resize::k_bilinear_fwd<<<resize::grid_for(planes * H * W), 256, 0, st>>>(
        x, planes, h, w, H, W, (float)h / (float)H, (float)w / (float)W, out));
    return OESS_OK;
}

// g [planes, H, W] -> dx [planes, h, w]; tmp: planes * H * w floats of scratch
OESS_API int oess_bilinear_resize_planes_bwd(const float* g, int64_t planes, int h, int w, int H, int W, float* tmp, float* dx,
                                             oess_stream_t stream) {
    if (!g || !tmp || !dx || planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("bilinear_resize_planes_bwd_x", st, resize::k_bilinear_bwd_x<<<resize::grid_for(planes * H * w), 256, 0, st>>>(
        g, planes * H, w, W, (float)w / (float)W, tmp));
    OESS_KERNEL("bilinear_resize_planes_bwd_y", st, resize::k_bilinear_bwd_y<<<resize::grid_for(planes * h * w), 256, 0, st>>>(
        tmp, planes, h, w, H, (float)h / (float)H, dx));
    return OESS_OK;
}
