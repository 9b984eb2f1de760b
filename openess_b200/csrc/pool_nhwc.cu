// pool_nhwc.cu -- the two pooling layers of the torchvision-style ResNets on the path (models/_resnet.py:137 MaxPool2d(3, 2, 1)
// after the stem, :149 AdaptiveAvgPool2d((1, 1)) before fc) on channels-last float32 tensors.  HBM-bound, no reuse beyond
// the 3x3 window (served by L1/L2): 4 B read + 1 B written per input element for the max pool.
#include <float.h>

#include "common.cuh"

namespace oess {
namespace pool {

// thread = one output pixel x 4 channels
__global__ void __launch_bounds__(256)
k_maxpool3x3s2(const float4* __restrict__ x, int B, int H, int W, int C4, int Ho, int Wo, float4* __restrict__ y) {
    const int64_t total = (int64_t)B * Ho * Wo * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int xo = (int)((i / C4) % Wo);
        const int yo = (int)((i / ((int64_t)C4 * Wo)) % Ho);
        const int b = (int)(i / ((int64_t)C4 * Wo * Ho));
        float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);      // padding never wins (torch pads with -inf)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int yi = 2 * yo - 1 + dy;
            if (yi < 0 || yi >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int xi = 2 * xo - 1 + dx;
                if (xi < 0 || xi >= W) continue;
                const float4 v = x[(((int64_t)b * H + yi) * W + xi) * C4 + c];
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        y[i] = m;
    }
}

// grid (ceil(C / 32), B), 256 threads = 8 pixel-slices x 32 channels; fp32 partial sums, one shared-memory combine
__global__ void __launch_bounds__(256)
k_global_avgpool(const float* __restrict__ x, int64_t HW, int C, float* __restrict__ y) {
    __shared__ float s[8][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const float* xb = x + (int64_t)blockIdx.y * HW * C;
    float acc = 0.0f;
    if (c < C)
        for (int64_t p = slice; p < HW; p += 8) acc += xb[p * C + c];
    s[slice][lane] = acc;
    __syncthreads();
    if (slice == 0 && c < C) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s[k][lane];
        y[(int64_t)blockIdx.y * C + c] = t / (float)HW;
    }
}

}  // namespace pool
}  // namespace oess

using namespace oess;

OESS_API int oess_maxpool3x3s2_nhwc(const float* x, int B, int H, int W, int C, float* y, oess_stream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!x || !y || (((uintptr_t)x | (uintptr_t)y) & 15)) return OESS_E_ARG;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("maxpool3x3s2_nhwc", st, pool::k_maxpool3x3s2<<<(unsigned)g, 256, 0, st>>>((const float4*)x, B, H, W, C / 4, Ho, Wo, (float4*)y));
    return OESS_OK;
}

OESS_API int oess_global_avgpool_nhwc(const float* x, int B, int64_t HW, int C, float* y, oess_stream_t stream) {
    if (B < 0 || HW <= 0 || C <= 0 || B > 65535) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!x || !y) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("global_avgpool_nhwc", st, pool::k_global_avgpool<<<dim3((unsigned)((C + 31) / 32), (unsigned)B), 256, 0, st>>>(x, HW, C, y));
    return OESS_OK;
}

// ---- E2VID decoder helpers (online reconstruction; e2vid/model/unet.py:165-168, submodules.py:34-63) ----
namespace oess {
namespace pool {

// z[b, 2y + dy, 2x + dx, c] = (dy == 0 && dx == 0) ? x[b, y, x, c] + skip[b, y, x, c] : 0 -- the zero-inserted input that
// turns ConvTranspose2d(k, stride 2, padding p, output_padding 1) into a stride-1 convolution with the rotated kernel and
// padding k - 1 - p (the trailing zero row / column of z is the output_padding).  The 'sum' skip connection is fused.
__global__ void __launch_bounds__(256)
k_zero_insert2x(const float4* __restrict__ x, const float4* __restrict__ skip, int B, int H, int W, int C4, float4* __restrict__ z) {
    const int64_t total = (int64_t)B * 2 * H * 2 * W * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int xo = (int)((i / C4) % (2 * W));
        const int yo = (int)((i / ((int64_t)C4 * 2 * W)) % (2 * H));
        const int b = (int)(i / ((int64_t)C4 * 2 * W * 2 * H));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((xo | yo) & 1) == 0) {
            const int64_t src = (((int64_t)b * H + (yo >> 1)) * W + (xo >> 1)) * C4 + c;
            v = x[src];
            if (skip) {
                const float4 s = skip[src];
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
            }
        }
        z[i] = v;
    }
}

// out [B, 2H, 2W, C1 + C2] = cat(nearest-x2(x [B, H, W, C1]), skip [B, 2H, 2W, C2]) along the channels: the decoder transitions of
// SemSegE2VID (models/style_networks.py:148-158: f.interpolate(scale_factor=2, mode='nearest') + torch.cat) in one pass.
__global__ void __launch_bounds__(256)
k_up2x_cat(const float4* __restrict__ x, const float4* __restrict__ skip, int B, int H, int W, int C1q, int C2q,
           float4* __restrict__ out) {
    const int Cq = C1q + C2q;
    const int64_t total = (int64_t)B * 2 * H * 2 * W * Cq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cq);
        const int64_t pix = i / Cq;
        if (c < C1q) {
            const int xo = (int)(pix % (2 * W));
            const int yo = (int)((pix / (2 * W)) % (2 * H));
            const int b = (int)(pix / ((int64_t)4 * W * H));
            out[i] = __ldg(x + (((int64_t)b * H + (yo >> 1)) * W + (xo >> 1)) * C1q + c);
        } else {
            out[i] = __ldcs(skip + pix * C2q + (c - C1q));
        }
    }
}
// backward: dx [B, H, W, C1] = 2 x 2 sums of g[..., :C1]; dskip [B, 2H, 2W, C2] = g[..., C1:] (either may be NULL)
__global__ void __launch_bounds__(256)
k_up2x_cat_bwd(const float4* __restrict__ g, int B, int H, int W, int C1q, int C2q, float4* __restrict__ dx,
               float4* __restrict__ dskip) {
    const int Cq = C1q + C2q;
    const int64_t n1 = dx ? (int64_t)B * H * W * C1q : 0, n2 = dskip ? (int64_t)B * 4 * H * W * C2q : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n1) {
            const int c = (int)(i % C1q);
            const int xi = (int)((i / C1q) % W);
            const int yi = (int)((i / ((int64_t)C1q * W)) % H);
            const int b = (int)(i / ((int64_t)C1q * W * H));
            const int64_t row = ((int64_t)b * 2 * H + 2 * yi) * (2 * W) + 2 * xi;
            const float4 a0 = __ldcs(g + row * Cq + c), a1 = __ldcs(g + (row + 1) * Cq + c);
            const float4 a2 = __ldcs(g + (row + 2 * W) * Cq + c), a3 = __ldcs(g + (row + 2 * W + 1) * Cq + c);
            dx[i] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                                (a0.w + a1.w) + (a2.w + a3.w));
        } else {
            const int64_t j = i - n1;
            dskip[j] = __ldcs(g + (j / C2q) * Cq + C1q + (int)(j % C2q));
        }
    }
}

// out[p] = sigmoid(sum_c w[c] * (x[p, c] + skip[p, c]) + bias): the prediction layer (1x1 conv to one channel + sigmoid) with
// the last skip sum fused.  One warp per 32 / (C / 4)... simple form: thread per pixel, C <= 64.
__global__ void __launch_bounds__(256)
k_pred_sigmoid(const float4* __restrict__ x, const float4* __restrict__ skip, const float* __restrict__ w, float bias, int64_t P,
               int C4, float* __restrict__ out) {
    __shared__ float4 s_w[16];
    if (threadIdx.x < C4) s_w[threadIdx.x] = reinterpret_cast<const float4*>(w)[threadIdx.x];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        float acc = bias;
        for (int c = 0; c < C4; ++c) {
            float4 v = x[p * C4 + c];
            if (skip) {
                const float4 s = skip[p * C4 + c];
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
            }
            const float4 ww = s_w[c];
            acc += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
        }
        out[p] = 1.0f / (1.0f + __expf(-acc));
    }
}

}  // namespace pool
}  // namespace oess

OESS_API int oess_zero_insert2x_nhwc(const float* x, const float* skip, int B, int H, int W, int C, float* z, oess_stream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!x || !z || (((uintptr_t)x | (uintptr_t)skip | (uintptr_t)z) & 15)) return OESS_E_ARG;
    const int64_t total = (int64_t)B * 4 * H * W * (C / 4);
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("zero_insert2x_nhwc", st, pool::k_zero_insert2x<<<(unsigned)g, 256, 0, st>>>((const float4*)x, (const float4*)skip, B, H, W, C / 4, (float4*)z));
    return OESS_OK;
}

OESS_API int oess_upsample2x_cat_nhwc(const float* x, const float* skip, int B, int H, int W, int C1, int C2, float* out,
                                      oess_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C1 <= 0 || C2 < 0 || (C1 & 3) || (C2 & 3)) return OESS_E_ARG;
    if (!x || !out || (C2 > 0 && !skip) || (((uintptr_t)x | (uintptr_t)skip | (uintptr_t)out) & 15)) return OESS_E_ARG;
    const int64_t total = (int64_t)B * 4 * H * W * ((C1 + C2) / 4);
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("upsample2x_cat_nhwc", st, pool::k_up2x_cat<<<(unsigned)g, 256, 0, st>>>(
        (const float4*)x, (const float4*)skip, B, H, W, C1 / 4, C2 / 4, (float4*)out));
    return OESS_OK;
}

OESS_API int oess_upsample2x_cat_nhwc_bwd(const float* g, int B, int H, int W, int C1, int C2, float* dx, float* dskip,
                                          oess_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C1 <= 0 || C2 < 0 || (C1 & 3) || (C2 & 3)) return OESS_E_ARG;
    if (!g || (!dx && !dskip) || (dskip && C2 == 0) || (((uintptr_t)g | (uintptr_t)dx | (uintptr_t)dskip) & 15)) return OESS_E_ARG;
    const int64_t total = (dx ? (int64_t)B * H * W * (C1 / 4) : 0) + (dskip ? (int64_t)B * 4 * H * W * (C2 / 4) : 0);
    int64_t gr = (total + 255) / 256;
    if (gr > (int64_t)kNumSMs * 16) gr = (int64_t)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("upsample2x_cat_nhwc_bwd", st, pool::k_up2x_cat_bwd<<<(unsigned)gr, 256, 0, st>>>(
        (const float4*)g, B, H, W, C1 / 4, C2 / 4, (float4*)dx, (float4*)dskip));
    return OESS_OK;
}

OESS_API int oess_pred_sigmoid_nhwc(const float* x, const float* skip, const float* w, float bias, int64_t pixels, int C,
                                    float* out, oess_stream_t stream) {
    if (pixels < 0 || C <= 0 || (C & 3) || C > 64) return OESS_E_ARG;
    if (pixels == 0) return OESS_OK;
    if (!x || !w || !out || (((uintptr_t)x | (uintptr_t)skip | (uintptr_t)w) & 15)) return OESS_E_ARG;
    int64_t g = (pixels + 255) / 256;
    if (g > (int64_t)kNumSMs * 16) g = (int64_t)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("pred_sigmoid_nhwc", st, pool::k_pred_sigmoid<<<(unsigned)g, 256, 0, st>>>((const float4*)x, (const float4*)skip, w, bias, pixels, C / 4, out));
    return OESS_OK;
}
