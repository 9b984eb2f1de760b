// pointwise.cu -- HBM-bound pointwise / reduce kernels that sit between the dense contractions of the path.
//   a10  e2vid/model/submodules.py:197-214  ConvLSTM gate non-linearities + cell/hidden update, fused in one pass
//        (the reference runs ~10 ATen elementwise kernels per encoder level per recurrent step)
//   a18  training/openess_trainer.py:456-462  consistency losses: L1Loss(feat_a, feat_b) and
//        mean(1 - cosine_similarity(logits_a, logits_b, dim=1))
#include "common.cuh"

namespace oess {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// gates [B, 4C, HW] = (in, remember, out, cell) chunks; prev_cell [B, C, HW] or NULL (zeros)
// cell = remember * prev_cell + in * tanh(cell_gate); hidden = out * tanh(cell)     (submodules.py:203-212)
__global__ void __launch_bounds__(256)
k_convlstm_gates(const float* __restrict__ gates, const float* __restrict__ prev_cell, float* __restrict__ hidden,
                 float* __restrict__ cell, int64_t CHW, int64_t total4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total4; v += stride) {
        const int64_t e = v * 4;                  // element index in [B, C, HW]; CHW % 4 == 0
        const int64_t b = e / CHW, r = e - b * CHW;
        const float* g = gates + b * 4 * CHW + r;
        const float4 gi = __ldcs(reinterpret_cast<const float4*>(g));
        const float4 gr = __ldcs(reinterpret_cast<const float4*>(g + CHW));
        const float4 go = __ldcs(reinterpret_cast<const float4*>(g + 2 * CHW));
        const float4 gc = __ldcs(reinterpret_cast<const float4*>(g + 3 * CHW));
        float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_cell) pc = __ldcs(reinterpret_cast<const float4*>(prev_cell + e));
        float4 c, h;
        c.x = sigmoidf_(gr.x) * pc.x + sigmoidf_(gi.x) * tanhf(gc.x);
        c.y = sigmoidf_(gr.y) * pc.y + sigmoidf_(gi.y) * tanhf(gc.y);
        c.z = sigmoidf_(gr.z) * pc.z + sigmoidf_(gi.z) * tanhf(gc.z);
        c.w = sigmoidf_(gr.w) * pc.w + sigmoidf_(gi.w) * tanhf(gc.w);
        h.x = sigmoidf_(go.x) * tanhf(c.x);
        h.y = sigmoidf_(go.y) * tanhf(c.y);
        h.z = sigmoidf_(go.z) * tanhf(c.z);
        h.w = sigmoidf_(go.w) * tanhf(c.w);
        *reinterpret_cast<float4*>(cell + e) = c;
        *reinterpret_cast<float4*>(hidden + e) = h;
    }
}

__global__ void __launch_bounds__(256)
k_convlstm_gates_scalar(const float* __restrict__ gates, const float* __restrict__ prev_cell,
                        float* __restrict__ hidden, float* __restrict__ cell, int64_t CHW, int64_t total) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t b = e / CHW, r = e - b * CHW;
        const float* g = gates + b * 4 * CHW + r;
        const float pc = prev_cell ? prev_cell[e] : 0.f;
        const float c = sigmoidf_(g[CHW]) * pc + sigmoidf_(g[0]) * tanhf(g[3 * CHW]);
        cell[e] = c;
        hidden[e] = sigmoidf_(g[2 * CHW]) * tanhf(c);
    }
}

// block-level float64 sum -> one atomicAdd(double)
__device__ __forceinline__ void block_accumulate(double v, double* dst) {
    __shared__ double s_red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += s_red[k];
        atomicAdd(dst, a);
    }
}

// sum |a - b| -> acc[0] (float64)
__global__ void __launch_bounds__(256)
k_l1_sum(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        s += fabsf(__ldcs(a + i) - __ldcs(b + i));
    block_accumulate((double)s, acc);
}

// d mean|a-b| / da = sign(a - b) / n * g;  db = -da
__global__ void __launch_bounds__(256)
k_l1_bwd(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ gscale,
         float* __restrict__ da, float* __restrict__ db) {
    const float g = gscale[0] / (float)n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = __ldcs(a + i) - __ldcs(b + i);
        const float s = (d > 0.f) ? g : ((d < 0.f) ? -g : 0.f);      // torch: sign(0) = 0
        if (da) __stcs(da + i, s);
        if (db) __stcs(db + i, -s);
    }
}

// sum over pixels of (1 - cos(a_px, b_px)) with the channel dim strided by HW (NCHW); torch semantics:
// cos = sum_c (a_c / max(|a|, eps)) * (b_c / max(|b|, eps)), eps = 1e-8
template <bool BWD>
__global__ void __launch_bounds__(256)
k_cos_consistency(const float* __restrict__ a, const float* __restrict__ b, int B, int K, int64_t HW,
                  double* __restrict__ acc, const float* __restrict__ gscale, float* __restrict__ da,
                  float* __restrict__ db) {
    const int64_t total = (int64_t)B * HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float eps = 1e-8f;
    float s = 0.f;
    const float g = BWD ? -gscale[0] / (float)total : 0.f;          // d mean(1 - cos) = -dcos / total
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t bb = i / HW, px = i - bb * HW;
        const float* pa = a + bb * K * HW + px;
        const float* pb = b + bb * K * HW + px;
        float ab = 0.f, aa = 0.f, bq = 0.f;
        for (int c = 0; c < K; ++c) {
            const float x = pa[(int64_t)c * HW], y = pb[(int64_t)c * HW];
            ab += x * y; aa += x * x; bq += y * y;
        }
        const float na = fmaxf(sqrtf(aa), eps), nb = fmaxf(sqrtf(bq), eps);
        const float cosv = ab / (na * nb);
        if (!BWD) {
            s += 1.0f - cosv;
        } else {
            // dcos/da_c = b_c/(na nb) - cos * a_c / na^2   (for |a| > eps; the clamp region has zero measure)
            const float inv = 1.0f / (na * nb), ca = cosv / (na * na), cb = cosv / (nb * nb);
            for (int c = 0; c < K; ++c) {
                const float x = pa[(int64_t)c * HW], y = pb[(int64_t)c * HW];
                if (da) __stcs(da + bb * K * HW + (int64_t)c * HW + px, g * (y * inv - ca * x));
                if (db) __stcs(db + bb * K * HW + (int64_t)c * HW + px, g * (x * inv - cb * y));
            }
        }
    }
    if (!BWD) block_accumulate((double)s, acc);
}

__global__ void k_finish_mean(const double* __restrict__ acc, double n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(acc[0] / n);
}

static inline unsigned ew_grid(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace oess

using namespace oess;

OESS_API int oess_convlstm_gates(const float* gates, const float* prev_cell, float* hidden, float* cell, int B, int C,
                                 int64_t HW, oess_stream_t stream) {
    if (B <= 0 || C <= 0 || HW <= 0 || !gates || !hidden || !cell) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t CHW = (int64_t)C * HW, total = (int64_t)B * CHW;
    const bool vec = (CHW % 4 == 0) && (((uintptr_t)gates | (uintptr_t)hidden | (uintptr_t)cell | (uintptr_t)prev_cell) % 16 == 0);
    if (vec) {
        OESS_KERNEL("convlstm_gates", st, k_convlstm_gates<<<ew_grid(total / 4), 256, 0, st>>>(gates, prev_cell, hidden, cell, CHW, total / 4));
    } else {
        OESS_KERNEL("convlstm_gates", st, k_convlstm_gates_scalar<<<ew_grid(total), 256, 0, st>>>(gates, prev_cell, hidden, cell, CHW, total));
    }
    return OESS_OK;
}

OESS_API int oess_l1_mean(const float* a, const float* b, int64_t n, float* loss, double* acc, oess_stream_t stream) {
    if (n <= 0 || !a || !b || !loss || !acc) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
    OESS_KERNEL("l1_sum", st, k_l1_sum<<<ew_grid(n / 4 + 1), 256, 0, st>>>(a, b, n, acc));
    OESS_KERNEL("finish_mean", st, k_finish_mean<<<1, 32, 0, st>>>(acc, (double)n, loss));
    return OESS_OK;
}

OESS_API int oess_l1_mean_bwd(const float* a, const float* b, int64_t n, const float* grad_scale, float* da, float* db,
                              oess_stream_t stream) {
    if (n <= 0 || !a || !b || !grad_scale || (!da && !db)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("l1_bwd", st, k_l1_bwd<<<ew_grid(n / 2 + 1), 256, 0, st>>>(a, b, n, grad_scale, da, db));
    return OESS_OK;
}

OESS_API int oess_cos_consistency(const float* a, const float* b, int B, int K, int64_t HW, float* loss, double* acc,
                                  oess_stream_t stream) {
    if (B <= 0 || K <= 0 || HW <= 0 || !a || !b || !loss || !acc) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
    OESS_KERNEL("cos_consistency", st, k_cos_consistency<false><<<ew_grid((int64_t)B * HW), 256, 0, st>>>(
        a, b, B, K, HW, acc, nullptr, nullptr, nullptr));
    OESS_KERNEL("finish_mean", st, k_finish_mean<<<1, 32, 0, st>>>(acc, (double)B * (double)HW, loss));
    return OESS_OK;
}

OESS_API int oess_cos_consistency_bwd(const float* a, const float* b, int B, int K, int64_t HW, const float* grad_scale,
                                      float* da, float* db, oess_stream_t stream) {
    if (B <= 0 || K <= 0 || HW <= 0 || !a || !b || !grad_scale || (!da && !db)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("cos_consistency_bwd", st, k_cos_consistency<true><<<ew_grid((int64_t)B * HW), 256, 0, st>>>(
        a, b, B, K, HW, nullptr, grad_scale, da, db));
    return OESS_OK;
}
