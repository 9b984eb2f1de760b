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
