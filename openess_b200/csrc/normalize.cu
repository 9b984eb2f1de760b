// normalize.cu -- nonzero mean/std standardisation, HBM-bound two-pass (stats, apply).
//   a4  datasets/data_util.py:38-48            normalize_voxel_grid        (biased, whole tensor)
//   a8  e2vid/utils/inference_utils.py:77-85   EventPreprocessor.__call__  (biased, whole batch tensor)
//       DSEC/dataset/representations.py:45-53  VoxelGrid(normalize=True)   (unbiased torch.std, std > 0 guard)
// Algorithmic bytes: 2 reads + 1 write of the tensor = 12 B / element.
// Sums are accumulated in float64 (per-thread -> warp shuffle -> one atomicAdd(double) per CTA), so
// the statistics are at least as accurate as torch's float32 reductions; parity is tolerance-based.
#include "normalize.cuh"

namespace oess {
namespace norm {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k_stats(const float* __restrict__ x, int64_t group_numel, double* __restrict__ stats) {
    const int g = blockIdx.y;
    const float* xg = x + (int64_t)g * group_numel;
    double s = 0.0, ss = 0.0;
    unsigned long long nz = 0;
    const int64_t stride = (int64_t)gridDim.x * kThreads * 4;
    int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(xg) & 15) == 0);
    for (; i < group_numel; i += stride) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec && i + 3 < group_numel) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(xg + i));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k < group_numel) v[k] = xg[i + k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (v[k] != 0.0f) { s += (double)v[k]; ss += (double)v[k] * (double)v[k]; ++nz; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
    }
    __shared__ double sh[3][kThreads / 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = (double)nz; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) a += sh[threadIdx.x][k];
        atomicAdd(&stats[(int64_t)g * 3 + threadIdx.x], a);
    }
}

template <bool UNBIASED>
__global__ void __launch_bounds__(kThreads)
k_apply(float* __restrict__ x, int64_t group_numel, const double* __restrict__ stats) {
    const int g = blockIdx.y;
    float* xg = x + (int64_t)g * group_numel;
    const double sum = stats[(int64_t)g * 3], sumsq = stats[(int64_t)g * 3 + 1], nnz = stats[(int64_t)g * 3 + 2];
    if (!(nnz > 0)) return;  // `if num_nonzeros > 0` / `if mask[0].size()[0] > 0`
    float mean, sd;
    if (UNBIASED) {
        mean = (float)(sum / nnz);
        sd = (float)sqrt((sumsq - sum * sum / nnz) / (nnz - 1.0));  // torch.std(); nnz == 1 -> NaN
    } else {
        // mean = events.sum() / nnz ; stddev = sqrt((events**2).sum() / nnz - mean**2)   all float32
        const float fn = (float)nnz;
        mean = __fdiv_rn((float)sum, fn);
        sd = __fsqrt_rn(__fsub_rn(__fdiv_rn((float)sumsq, fn), __fmul_rn(mean, mean)));
    }
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < group_numel; i += stride) {
        const float v = xg[i];
        float r;
        if (UNBIASED) {
            if (v == 0.0f) continue;  // only voxel_grid[mask] is rewritten
            r = (sd > 0.0f) ? __fdiv_rn(__fsub_rn(v, mean), sd) : __fsub_rn(v, mean);
        } else {
            const float m = (v != 0.0f) ? 1.0f : 0.0f;
            r = __fdiv_rn(__fmul_rn(m, __fsub_rn(v, mean)), sd);  // mask * (events - mean) / stddev
        }
        xg[i] = r;
    }
}

// x [B, C, HW] planes -> y [B, HW, Cp] channels-last with the channels zero-padded to Cp (a multiple of 4), optionally
// applying the EventPreprocessor normalisation (biased, inference_utils.py:77-85) from precomputed stats on the way:
// the input transform in front of the tensor-core head convolution of E2VID (a 5-channel plane tensor is too thin for
// a 16-byte-aligned TMA row).
__global__ void __launch_bounds__(kThreads)
k_to_nhwc_padded(const float* __restrict__ x, int C, int64_t HW, int64_t total, const double* __restrict__ stats, int Cp,
                 float* __restrict__ y, int W, int pad_w) {
    bool norm = false;
    float mean = 0.f, sd = 1.f;
    if (stats) {
        const double sum = stats[0], sumsq = stats[1], nnz = stats[2];
        if (nnz > 0) {
            const float fn = (float)nnz;
            mean = __fdiv_rn((float)sum, fn);
            sd = __fsqrt_rn(__fsub_rn(__fdiv_rn((float)sumsq, fn), __fmul_rn(mean, mean)));
            norm = true;
        }
    }
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
        const int64_t b = i / HW, px = i - b * HW;
        const float* xp = x + b * C * HW + px;
        // pad_w > 0: the output rows are W + 2 pad_w pixels wide with zero borders (written by k_zero_borders)
        const int64_t row = (b * HW + px) / W;
        float* yp = y + (pad_w ? (i + (2 * row + 1) * (int64_t)pad_w) : i) * Cp;
        for (int c0 = 0; c0 < Cp; c0 += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                float t = (c < C) ? __ldcs(xp + (int64_t)c * HW) : 0.f;
                if (norm && c < C) t = __fdiv_rn(__fmul_rn((t != 0.0f) ? 1.0f : 0.0f, __fsub_rn(t, mean)), sd);
                v[j] = t;
            }
            *reinterpret_cast<float4*>(yp + c0) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// zero the pad_w border pixels at both ends of every row of y [rows, W + 2 pad_w, Cp]
__global__ void k_zero_borders(float* __restrict__ y, int64_t rows, int W, int pad_w, int Cp) {
    const int64_t n = rows * 2 * pad_w * Cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / (2 * pad_w * Cp);
        const int j = (int)(i - r * 2 * pad_w * Cp);
        const int px = j / Cp, c = j - px * Cp;
        const int col = px < pad_w ? px : W + px;          // left border, then right border
        y[(r * (W + 2 * pad_w) + col) * Cp + c] = 0.0f;
    }
}

}  // namespace norm
}  // namespace oess

using namespace oess;

// Same as oess_planes_to_nhwc_padded with the rows additionally zero-padded by pad_w pixels at both ends:
// y [B, H, W + 2 pad_w, Cp] (the input layout of oess_conv2d_nhwc_tf32_rowunfold).
OESS_API int oess_planes_to_nhwc_padded_w(const float* x, int B, int C, int H, int W, const double* stats, int Cp, int pad_w,
                                          float* y, oess_stream_t stream) {
    if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || Cp < C || (Cp & 3) || pad_w <= 0 || ((uintptr_t)y & 15)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W, total = (int64_t)B * HW;
    int64_t blocks = (total + norm::kThreads - 1) / norm::kThreads;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    OESS_KERNEL("k_zero_borders", st, norm::k_zero_borders<<<kNumSMs, 256, 0, st>>>(y, (int64_t)B * H, W, pad_w, Cp));
    OESS_KERNEL("k_to_nhwc_padded", st, norm::k_to_nhwc_padded<<<(unsigned)blocks, norm::kThreads, 0, st>>>(
        x, C, HW, total, stats, Cp, y, W, pad_w));
    return OESS_OK;
}

OESS_API int oess_planes_to_nhwc_padded(const float* x, int B, int C, int64_t HW, const double* stats, int Cp, float* y,
                                        oess_stream_t stream) {
    if (!x || !y || B <= 0 || C <= 0 || HW <= 0 || Cp < C || (Cp & 3) || ((uintptr_t)y & 15)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)B * HW;
    int64_t blocks = (total + norm::kThreads - 1) / norm::kThreads;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    OESS_KERNEL("k_to_nhwc_padded", st, norm::k_to_nhwc_padded<<<(unsigned)blocks, norm::kThreads, 0, st>>>(
        x, C, HW, total, stats, Cp, y, 1, 0));
    return OESS_OK;
}

int launch_nonzero_standardize(float* x, int64_t group_numel, int n_groups, double* stats, int phase,
                               int unbiased, cudaStream_t st) {
    if (!x || !stats || group_numel <= 0 || n_groups <= 0 || phase < 0 || phase > 2) return OESS_E_ARG;
    if (n_groups > 65535) return OESS_E_RANGE;
    // enough CTAs to fill the GPU, but no more than the data needs
    int64_t per_group = (group_numel + norm::kThreads * 4 - 1) / (norm::kThreads * 4);
    int64_t want = (int64_t)kNumSMs * 8 / n_groups + 1;
    const unsigned bx = (unsigned)(per_group < want ? per_group : want);
    if (phase == 0 || phase == 1) {
        OESS_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 3 * (size_t)n_groups, st));
        OESS_KERNEL("norm_stats", st, norm::k_stats<<<dim3(bx, (unsigned)n_groups), norm::kThreads, 0, st>>>(x, group_numel, stats));
    }
    if (phase == 0 || phase == 2) {
        OESS_KERNEL("norm_apply", st, if (unbiased)
            norm::k_apply<true><<<dim3(bx * 4, (unsigned)n_groups), norm::kThreads, 0, st>>>(x, group_numel, stats);
        else
            norm::k_apply<false><<<dim3(bx * 4, (unsigned)n_groups), norm::kThreads, 0, st>>>(x, group_numel, stats));
    }
    return OESS_OK;
}

OESS_API int oess_nonzero_standardize(float* x, int64_t group_numel, int n_groups, double* stats, int phase,
                                      int unbiased, oess_stream_t stream) {
    return launch_nonzero_standardize(x, group_numel, n_groups, stats, phase, unbiased, (cudaStream_t)stream);
}
