// bn_nhwc.cu -- BatchNorm2d over channels-last activations, the way the OpenESS trainers really run the "frozen"
// ResNet-50 teacher: `.train()` is called on it every step (training/pretrain_trainer.py:370-371), so its BatchNorm
// layers normalise with BATCH statistics and update their running statistics although the weights are frozen
// (models/image_model.py:116-117 only clears requires_grad).  torch.nn.BatchNorm2d semantics (torch/nn/modules/
// batchnorm.py): y = (x - mean) / sqrt(var_biased + eps) * gamma + beta; running = (1 - m) running + m stat, with the
// UNBIASED variance for running_var.
//   k_bn_stats   : per-channel sum / sum of squares of x [R, C] (R = B H W rows), fp32 per-thread partials, fp64 combine
//   k_bn_finalize: scale / shift per channel (+ running-statistics update) on the device: no host synchronisation
//   k_bn_apply   : y = act(x * scale[c] + shift[c] (+ residual)), in place, 128-bit accesses
// All three are HBM-bound: 4 B / element read (stats), 8-12 B / element (apply).
#include <cuda_bf16.h>

#include "common.cuh"

namespace oess {

// grid.x = row blocks; block = 256 threads = (C4 = C / 4 channel quads) x (256 / C4 row lanes); C4 <= 256, C4 | 256
// for C4 > 256 (C = 2048: C4 = 512) the quad index also runs over blockIdx.y.
__global__ void __launch_bounds__(256)
k_bn_stats(const float* __restrict__ x, int64_t R, int C, int rows_per_block, double* __restrict__ sums) {
    const int C4 = C >> 2;
    const int qpb = min(C4, 256);                      // channel quads handled by this block
    const int lanes = 256 / qpb;                       // row lanes
    const int q = blockIdx.y * qpb + (threadIdx.x % qpb);
    const int rl = threadIdx.x / qpb;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(r0 + rows_per_block, R);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = r0 + rl;
    for (; r + 3 * (int64_t)lanes < r1; r += 4 * (int64_t)lanes) {        // 4 independent 128-bit loads in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(x + (r + (int64_t)u * lanes) * C) + q);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
            ss.x += v[u].x * v[u].x; ss.y += v[u].y * v[u].y; ss.z += v[u].z * v[u].z; ss.w += v[u].w * v[u].w;
        }
    }
    for (; r < r1; r += lanes) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + q);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        ss.x += v.x * v.x; ss.y += v.y * v.y; ss.z += v.z * v.z; ss.w += v.w * v.w;
    }
    __shared__ float4 sh_s[256], sh_ss[256];
    sh_s[threadIdx.x] = s;
    sh_ss[threadIdx.x] = ss;
    __syncthreads();
    if (rl == 0) {
        double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        for (int l = 0; l < lanes; ++l) {
            const float4 u = sh_s[l * qpb + threadIdx.x], w = sh_ss[l * qpb + threadIdx.x];
            a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
            b[0] += w.x; b[1] += w.y; b[2] += w.z; b[3] += w.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sums[q * 4 + j], a[j]);
            atomicAdd(&sums[C + q * 4 + j], b[j]);
        }
    }
}

__global__ void k_bn_finalize(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ running_mean, float* __restrict__ running_var, int C, double count,
                              float eps, float momentum, int training, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean, var;
    if (training) {
        mean = sums[c] / count;
        var = sums[C + c] / count - mean * mean;       // biased
        if (var < 0) var = 0;
        if (running_mean) {
            const double unbiased = count > 1 ? var * count / (count - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const double inv = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0;
    scale[c] = (float)(g * inv);
    shift[c] = (float)((beta ? (double)beta[c] : 0.0) - mean * g * inv);
}

// y_bf (optional): bfloat16 copy of the result, the operand of a kind::f16 consumer (oess_conv2d_nhwc_bf16); with
// write_f32 == 0 only that copy is written (x keeps the raw conv output).
__global__ void __launch_bounds__(256)
k_bn_apply(float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
           const float* __restrict__ residual, int64_t total4, int C4, int relu, float* __restrict__ y_out,
           __nv_bfloat16* __restrict__ y_bf, int write_f32) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int q = (int)(i % C4);
        float4 v = reinterpret_cast<float4*>(x)[i];
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + q);
        const float4 sf = __ldg(reinterpret_cast<const float4*>(shift) + q);
        v.x = v.x * sc.x + sf.x; v.y = v.y * sc.y + sf.y; v.z = v.z * sc.z + sf.z; v.w = v.w * sc.w + sf.w;
        if (residual) {
            const float4 r = __ldcs(reinterpret_cast<const float4*>(residual) + i);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (write_f32) reinterpret_cast<float4*>(y_out ? y_out : x)[i] = v;
        if (y_bf) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            reinterpret_cast<uint2*>(y_bf)[i] = pk;
        }
    }
}

// InstanceNorm2d (affine = False, no running statistics: torch.nn.InstanceNorm2d defaults, models/style_networks.py:257,274,277)
// from per-sample sums: y = (x - mean_bc) / sqrt(var_bc + eps) (+ residual) (ReLU), in place, channels-last.
__global__ void __launch_bounds__(256)
k_in_apply(float* __restrict__ x, const double* __restrict__ sums, const float* __restrict__ residual, int64_t HW, int C,
           int64_t total4, float eps, int relu, float* __restrict__ y_out) {
    const int C4 = C >> 2;
    const int64_t per_sample4 = HW * C4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int64_t b = i / per_sample4;
        const int q = (int)(i % C4);
        const double* sb = sums + b * 2 * C;
        float4 v = reinterpret_cast<float4*>(x)[i];
        float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = q * 4 + j;
            const double mean = sb[c] / (double)HW;
            double var = sb[C + c] / (double)HW - mean * mean;
            if (var < 0) var = 0;
            o[j] = (float)(((double)o[j] - mean) * rsqrt(var + (double)eps));
        }
        if (y_out) reinterpret_cast<float4*>(x)[i] = make_float4(o[0], o[1], o[2], o[3]);   // training: keep x_hat, y separately
        if (residual) {
            const float4 r = __ldcs(reinterpret_cast<const float4*>(residual) + i);
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        if (relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
        reinterpret_cast<float4*>(y_out ? y_out : x)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// InstanceNorm backward, pass 1: per (sample, channel) sums of g and g * x_hat, g = dy (* [y > 0] when the forward applied a
// ReLU).  grid (row blocks, channel-quad blocks, B); same thread layout as k_bn_stats.
__global__ void __launch_bounds__(256)
k_in_bwd_stats(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ xhat, int64_t HW, int C,
               int rows_per_block, double* __restrict__ sums, const double* __restrict__ fwd_sums, float eps, int raw) {
    const int C4 = C >> 2;
    const int qpb = min(C4, 256);
    const int lanes = 256 / qpb;
    const int q = blockIdx.y * qpb + (threadIdx.x % qpb);
    const int rl = threadIdx.x / qpb;
    const int b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, HW);
    const int64_t base = (int64_t)b * HW;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < C4) {
        float mu[4] = {0.f, 0.f, 0.f, 0.f}, iv[4] = {1.f, 1.f, 1.f, 1.f};     // raw: xhat holds the un-normalised conv output
        if (raw) {
            const double* fs = fwd_sums + (int64_t)b * 2 * C;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double mean = fs[q * 4 + j] / (double)HW;
                double var = fs[C + q * 4 + j] / (double)HW - mean * mean;
                if (var < 0) var = 0;
                mu[j] = (float)mean;
                iv[j] = (float)rsqrt(var + (double)eps);
            }
        }
        for (int64_t r = r0 + rl; r < r1; r += lanes) {
            const int64_t i = (base + r) * C4 + q;
            float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
            if (y) {
                const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + i);
                g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
            }
            float4 xh = __ldg(reinterpret_cast<const float4*>(xhat) + i);
            if (raw) { xh.x = (xh.x - mu[0]) * iv[0]; xh.y = (xh.y - mu[1]) * iv[1]; xh.z = (xh.z - mu[2]) * iv[2]; xh.w = (xh.w - mu[3]) * iv[3]; }
            s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
            ss.x += g.x * xh.x; ss.y += g.y * xh.y; ss.z += g.z * xh.z; ss.w += g.w * xh.w;
        }
    }
    __shared__ float4 sh_s[256], sh_ss[256];
    sh_s[threadIdx.x] = s;
    sh_ss[threadIdx.x] = ss;
    __syncthreads();
    if (rl == 0 && q < C4) {
        double a[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
        for (int l = 0; l < lanes; ++l) {
            const float4 u = sh_s[l * qpb + threadIdx.x], w = sh_ss[l * qpb + threadIdx.x];
            a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
            c[0] += w.x; c[1] += w.y; c[2] += w.z; c[3] += w.w;
        }
        double* sb = sums + (int64_t)b * 2 * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sb[q * 4 + j], a[j]);
            atomicAdd(&sb[C + q * 4 + j], c[j]);
        }
    }
}

// pass 2: dz = inv_std * (g - mean(g) - x_hat * mean(g x_hat)); optionally d_res = g (the residual branch's gradient)
__global__ void __launch_bounds__(256)
k_in_bwd_apply(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ xhat,
               const double* __restrict__ fwd_sums, const double* __restrict__ bwd_sums, int64_t HW, int C, int64_t total4,
               float eps, float* __restrict__ dz, float* __restrict__ d_res, const float* __restrict__ gamma, int raw) {
    const int C4 = C >> 2;
    const int64_t per_sample4 = HW * C4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int64_t b = i / per_sample4;
        const int q = (int)(i % C4);
        const double* fs = fwd_sums + b * 2 * C;
        const double* bs = bwd_sums + b * 2 * C;
        float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
        if (y) {
            const float4 yy = __ldg(reinterpret_cast<const float4*>(y) + i);
            g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
        }
        const float4 xh = __ldg(reinterpret_cast<const float4*>(xhat) + i);
        const float gg[4] = {g.x, g.y, g.z, g.w}, xx[4] = {xh.x, xh.y, xh.z, xh.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = q * 4 + j;
            const double mean = fs[c] / (double)HW;
            double var = fs[C + c] / (double)HW - mean * mean;
            if (var < 0) var = 0;
            const double inv = rsqrt(var + (double)eps);
            const double xhj = raw ? ((double)xx[j] - mean) * inv : (double)xx[j];
            const double gam = gamma ? (double)gamma[c] : 1.0;
            o[j] = (float)(gam * inv * ((double)gg[j] - bs[c] / (double)HW - xhj * (bs[C + c] / (double)HW)));
        }
        reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (d_res) reinterpret_cast<float4*>(d_res)[i] = g;
    }
}

}  // namespace oess

using namespace oess;

// x: [B, HW, C] channels-last, in place; sums: [B][2 C] doubles from oess_conv2d_nhwc_tf32_instats.
// Backward of y = act(InstanceNorm(z) + residual): dy, y (NULL when the forward had no ReLU), x_hat [B, HW, C] channels-last;
// fwd_sums = the forward's per-sample sums of z; bwd_sums [B][2 C] doubles of scratch (zeroed inside); dz (gradient w.r.t.
// the conv output z) and optionally d_res (gradient of the residual input) are written.
OESS_API int oess_instancenorm_nhwc_bwd(const float* dy, const float* y, const float* xhat, int B, int64_t HW, int C,
                                        const double* fwd_sums, double* bwd_sums, float eps, float* dz, float* d_res,
                                        oess_stream_t stream) {
    if (!dy || !xhat || !fwd_sums || !bwd_sums || !dz || B <= 0 || HW <= 0 || C <= 0 || (C & 3) || B > 65535) return OESS_E_ARG;
    const int C4 = C >> 2;
    if (C4 > 256 ? (C4 % 256) != 0 : (256 % C4) != 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(bwd_sums, 0, sizeof(double) * 2 * (size_t)C * B, st));
    const int qpb = C4 < 256 ? C4 : 256;
    const int lanes = 256 / qpb;
    int64_t rpb = (HW * B + (int64_t)kNumSMs * 16 - 1) / ((int64_t)kNumSMs * 16);
    if (rpb < (int64_t)lanes * 16) rpb = (int64_t)lanes * 16;
    const dim3 grid((unsigned)((HW + rpb - 1) / rpb), (unsigned)((C4 + qpb - 1) / qpb), (unsigned)B);
    OESS_KERNEL("in_bwd_stats", st, k_in_bwd_stats<<<grid, 256, 0, st>>>(dy, y, xhat, HW, C, (int)rpb, bwd_sums, fwd_sums, eps, 0));
    const int64_t total4 = (int64_t)B * HW * C4;
    const unsigned blocks = (unsigned)((total4 + 255) / 256 < (int64_t)kNumSMs * 16 ? (total4 + 255) / 256 : (int64_t)kNumSMs * 16);
    OESS_KERNEL("in_bwd_apply", st, k_in_bwd_apply<<<blocks, 256, 0, st>>>(dy, y, xhat, fwd_sums, bwd_sums, HW, C, total4, eps, dz, d_res, nullptr, 0));
    return OESS_OK;
}

// Training forward: x is overwritten with x_hat (kept for the backward pass), y_out receives act(x_hat + residual).
OESS_API int oess_instancenorm_nhwc_sums_train(float* x, int B, int64_t HW, int C, const double* sums, float eps,
                                               const float* residual, int relu, float* y_out, oess_stream_t stream) {
    if (!x || !sums || !y_out || B <= 0 || HW <= 0 || C <= 0 || (C & 3)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total4 = (int64_t)B * HW * (C >> 2);
    const unsigned blocks = (unsigned)((total4 + 255) / 256 < (int64_t)kNumSMs * 16 ? (total4 + 255) / 256 : (int64_t)kNumSMs * 16);
    OESS_KERNEL("in_apply", st, k_in_apply<<<blocks, 256, 0, st>>>(x, sums, residual, HW, C, total4, eps, relu ? 1 : 0, y_out));
    return OESS_OK;
}

OESS_API int oess_instancenorm_nhwc_sums(float* x, int B, int64_t HW, int C, const double* sums, float eps,
                                         const float* residual, int relu, oess_stream_t stream) {
    if (!x || !sums || B <= 0 || HW <= 0 || C <= 0 || (C & 3)) return OESS_E_ARG;
    if (((uintptr_t)x | (uintptr_t)residual) & 15) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total4 = (int64_t)B * HW * (C >> 2);
    const unsigned blocks = (unsigned)((total4 + 255) / 256 < (int64_t)kNumSMs * 16 ? (total4 + 255) / 256 : (int64_t)kNumSMs * 16);
    OESS_KERNEL("in_apply", st, k_in_apply<<<blocks, 256, 0, st>>>(x, sums, residual, HW, C, total4, eps, relu ? 1 : 0, nullptr));
    return OESS_OK;
}

// x: [R, C] channels-last rows (R = B * H * W), normalised IN PLACE.  ws: 2 C doubles + 2 C floats of device scratch
// (oess_bn_ws_bytes).  training != 0: batch statistics (+ running-statistics update when running_mean != NULL);
// training == 0: running statistics.  residual (same shape) is added after the affine map, before the ReLU.
OESS_API int oess_bn_ws_bytes(int C, size_t* ws_bytes) {
    if (!ws_bytes || C <= 0) return OESS_E_ARG;
    *ws_bytes = align_up(sizeof(double) * 2 * (size_t)C, 256) + align_up(sizeof(float) * 2 * (size_t)C, 256);
    return OESS_OK;
}

static int batchnorm_impl(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float eps, float momentum, int training, const float* residual,
                          int relu, void* ws, size_t ws_bytes, int have_sums, oess_stream_t stream, float* y_out = nullptr,
                          __nv_bfloat16* y_bf = nullptr, int write_f32 = 1) {
    size_t need = 0;
    if (oess_bn_ws_bytes(C, &need)) return OESS_E_ARG;
    if (!x || R < 0 || (C & 3)) return OESS_E_ARG;
    if (!training && (!running_mean || !running_var)) return OESS_E_ARG;
    if (!ws || ws_bytes < need) return OESS_E_WORKSPACE;
    if (((uintptr_t)x | (uintptr_t)residual | (uintptr_t)ws) & 15) return OESS_E_ARG;
    if (R == 0) return OESS_OK;
    const int C4 = C >> 2;
    if (C4 > 256 ? (C4 % 256) != 0 : (256 % C4) != 0) return OESS_E_ARG;     // C in {4, 8, .., 1024, 2048, ...}
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)ws;
    float* scale = (float*)((char*)ws + align_up(sizeof(double) * 2 * (size_t)C, 256));
    float* shift = scale + C;
    if (training && !have_sums) {
        OESS_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st));
        const int qpb = C4 < 256 ? C4 : 256;
        const int lanes = 256 / qpb;
        // ~16 CTAs per SM; every row lane walks >= 32 rows (8 rounds of 4 independent loads)
        int64_t rpb = (R + (int64_t)kNumSMs * 16 - 1) / ((int64_t)kNumSMs * 16);
        if (rpb < (int64_t)lanes * 32) rpb = (int64_t)lanes * 32;
        const dim3 grid((unsigned)((R + rpb - 1) / rpb), (unsigned)((C4 + qpb - 1) / qpb));
        OESS_KERNEL("bn_stats", st, k_bn_stats<<<grid, 256, 0, st>>>(x, R, C, (int)rpb, sums));
    }
    OESS_KERNEL("bn_finalize", st, k_bn_finalize<<<(C + 127) / 128, 128, 0, st>>>(
        sums, gamma, beta, running_mean, running_var, C, (double)R, eps, momentum, training ? 1 : 0, scale, shift));
    const int64_t total4 = R * C4;
    const unsigned blocks = (unsigned)((total4 + 255) / 256 < (int64_t)kNumSMs * 16 ? (total4 + 255) / 256 : (int64_t)kNumSMs * 16);
    OESS_KERNEL("bn_apply", st, k_bn_apply<<<blocks, 256, 0, st>>>(x, scale, shift, residual, total4, C4, relu ? 1 : 0, y_out,
                                                                y_bf, write_f32));
    return OESS_OK;
}

OESS_API int oess_batchnorm_nhwc(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float eps, float momentum, int training, const float* residual,
                                 int relu, void* ws, size_t ws_bytes, oess_stream_t stream) {
    return batchnorm_impl(x, R, C, gamma, beta, running_mean, running_var, eps, momentum, training, residual, relu, ws,
                          ws_bytes, 0, stream);
}

// Train-mode variant whose per-channel sums were already accumulated into the first 2 C doubles of `ws` by
// oess_conv2d_nhwc_tf32_stats (the statistics pass fused into the producing convolution's epilogue).
OESS_API int oess_batchnorm_nhwc_sums(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, float eps, float momentum, const float* residual, int relu,
                                      void* ws, size_t ws_bytes, oess_stream_t stream) {
    return batchnorm_impl(x, R, C, gamma, beta, running_mean, running_var, eps, momentum, 1, residual, relu, ws, ws_bytes,
                          1, stream);
}

// oess_batchnorm_nhwc_sums for a frozen network run with bfloat16 operands: the result is ALSO (write_f32 != 0, in place) or ONLY
// (write_f32 == 0: x keeps the raw conv output) stored as bfloat16 in y_bf16, the input of the next oess_conv2d_nhwc_bf16.
OESS_API int oess_batchnorm_nhwc_sums_bf16(float* x, int64_t R, int C, const float* gamma, const float* beta,
                                           float* running_mean, float* running_var, float eps, float momentum,
                                           const float* residual, int relu, void* y_bf16, int write_f32, void* ws,
                                           size_t ws_bytes, oess_stream_t stream) {
    if (!y_bf16 || ((uintptr_t)y_bf16 & 7)) return OESS_E_ARG;
    return batchnorm_impl(x, R, C, gamma, beta, running_mean, running_var, eps, momentum, 1, residual, relu, ws, ws_bytes,
                          1, stream, nullptr, (__nv_bfloat16*)y_bf16, write_f32 ? 1 : 0);
}

// Training forward of conv -> BatchNorm (batch statistics already in the first 2 C doubles of `ws`): z (the conv output) is
// KEPT for the backward pass, y_out receives act(BN(z) + residual); running statistics are updated.
OESS_API int oess_batchnorm_nhwc_sums_train(float* z, int64_t R, int C, const float* gamma, const float* beta,
                                            float* running_mean, float* running_var, float eps, float momentum,
                                            const float* residual, int relu, float* y_out, void* ws, size_t ws_bytes,
                                            oess_stream_t stream) {
    if (!y_out) return OESS_E_ARG;
    return batchnorm_impl(z, R, C, gamma, beta, running_mean, running_var, eps, momentum, 1, residual, relu, ws, ws_bytes, 1,
                          stream, y_out);
}

// Backward of y = act(gamma * (z - mean) / sqrt(var + eps) + beta + residual) with batch statistics (fwd_sums [2 C]):
//   g = dy * [y > 0];  d_beta = sum g;  d_gamma = sum g x_hat;  dz = gamma inv_std (g - mean(g) - x_hat mean(g x_hat))
// bwd_sums [2 C] doubles receive (d_beta, d_gamma); dz and optionally d_res (= g) are written.  y = NULL: no ReLU.
OESS_API int oess_batchnorm_nhwc_bwd(const float* dy, const float* y, const float* z, int64_t R, int C, const double* fwd_sums,
                                     const float* gamma, double* bwd_sums, float eps, float* dz, float* d_res,
                                     oess_stream_t stream) {
    if (!dy || !z || !fwd_sums || !bwd_sums || !dz || R <= 0 || C <= 0 || (C & 3)) return OESS_E_ARG;
    const int C4 = C >> 2;
    if (C4 > 256 ? (C4 % 256) != 0 : (256 % C4) != 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(bwd_sums, 0, sizeof(double) * 2 * (size_t)C, st));
    const int qpb = C4 < 256 ? C4 : 256;
    const int lanes = 256 / qpb;
    int64_t rpb = (R + (int64_t)kNumSMs * 16 - 1) / ((int64_t)kNumSMs * 16);
    if (rpb < (int64_t)lanes * 16) rpb = (int64_t)lanes * 16;
    const dim3 grid((unsigned)((R + rpb - 1) / rpb), (unsigned)((C4 + qpb - 1) / qpb), 1);
    OESS_KERNEL("bn_bwd_stats", st, k_in_bwd_stats<<<grid, 256, 0, st>>>(dy, y, z, R, C, (int)rpb, bwd_sums, fwd_sums, eps, 1));
    const int64_t total4 = R * C4;
    const unsigned blocks = (unsigned)((total4 + 255) / 256 < (int64_t)kNumSMs * 16 ? (total4 + 255) / 256 : (int64_t)kNumSMs * 16);
    OESS_KERNEL("bn_bwd_apply", st, k_in_bwd_apply<<<blocks, 256, 0, st>>>(dy, y, z, fwd_sums, bwd_sums, R, C, total4, eps, dz, d_res, gamma, 1));
    return OESS_OK;
}
