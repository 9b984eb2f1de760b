// losses.cu -- loss-side kernels of the OpenESS pretrain step.
//   a15  training/pretrain_trainer.py:445-465   superpixel mean-pool (sparse one-hot matmul)  -> segpool
//   a16  utils/loss_functions.py:138-153        NCELoss (InfoNCE, temperature)                -> infonce
//   a17  utils/loss_functions.py:6-24,96-135    TaskLoss = DiceLoss + CrossEntropyLoss(ignore)-> dice_ce
//   a20  evaluation/metrics.py:4-23             confusion matrix                               -> confusion
// All float32 inputs/outputs like the reference; cross-thread reductions are accumulated in float64 so
// the results are at least as accurate as torch's float32 reductions (parity is tolerance-based;
// confusion is integer-exact).
#include "common.cuh"

namespace oess {

// =============================================================================================
// a20 confusion matrix.  HBM-bound: 16 B / pixel.  Per-CTA smem histogram, 64-bit flush.
// =============================================================================================
constexpr int kConfSmemBins = 4096;  // K <= 64

__global__ void __launch_bounds__(256)
k_confusion(const int64_t* __restrict__ pred, const int64_t* __restrict__ gt, int64_t n, int K, int64_t ignore,
            unsigned long long* __restrict__ conf, int32_t* __restrict__ status) {
    __shared__ unsigned int s_bins[kConfSmemBins];
    const int KK = K * K;
    const bool use_smem = KK <= kConfSmemBins;
    if (use_smem) {
        for (int i = threadIdx.x; i < KK; i += blockDim.x) s_bins[i] = 0;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t g = __ldcs(gt + i);
        if (g == ignore) continue;                           // metrics.py:15-17
        const int64_t v = __ldcs(pred + i) + (int64_t)K * g; // :19
        if (v < 0 || v >= KK) { if (status) *status = 1; continue; }   // bincount would violate the assert :21
        if (use_smem) atomicAdd(&s_bins[v], 1u); else atomicAdd(&conf[v], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < KK; i += blockDim.x)
            if (s_bins[i]) atomicAdd(&conf[i], (unsigned long long)s_bins[i]);
    }
}

// Validation loop fused: argmax over the K logits of a pixel + confusion update in one pass; the int64
// prediction map of base_trainer_ov.py:463-466 (`pred.argmax(dim=1)`) is never materialised.  torch.argmax semantics:
// first index of the maximum, NaN counts as the maximum.
__global__ void __launch_bounds__(256)
k_argmax_confusion(const float* __restrict__ logits, const int64_t* __restrict__ gt, int B, int K, int64_t HW,
                   int64_t ignore, unsigned long long* __restrict__ conf, int32_t* __restrict__ status) {
    __shared__ unsigned int s_bins[kConfSmemBins];
    const int KK = K * K;
    for (int i = threadIdx.x; i < KK; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
    const int64_t total = (int64_t)B * HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t g = __ldcs(gt + i);
        if (g == ignore) continue;
        if (g < 0 || g >= K) { if (status) *status = 1; continue; }
        const int64_t b = i / HW, px = i - b * HW;
        const float* lp = logits + (b * K) * HW + px;
        float best = __ldcs(lp);
        int arg = 0;
        for (int c = 1; c < K; ++c) {
            const float v = __ldcs(lp + (int64_t)c * HW);
            if ((v > best && best == best) || (v != v && best == best)) { best = v; arg = c; }
        }
        atomicAdd(&s_bins[(int)g * K + arg], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KK; i += blockDim.x)
        if (s_bins[i]) atomicAdd(&conf[i], (unsigned long long)s_bins[i]);
}

// =============================================================================================
// a17 Dice + CE.  One pass over the logits (4*K + 8 B / pixel), per-thread register partials,
// block reduction in float64, one atomicAdd(double) per class per CTA.
//   partials[0..K)   inter_c = sum p_c * t_c         (over valid pixels)
//   partials[K..2K)  denom_c = sum p_c^2 + t_c^2
//   partials[2K]     ce_sum  = sum (lse - logit_target)
//   partials[2K+1]   n_valid
// =============================================================================================
template <int KMAX>
__global__ void __launch_bounds__(256)
k_dice_ce_partials(const float* __restrict__ logits, const int64_t* __restrict__ target, int B, int K, int64_t HW,
                   int64_t ignore, double* __restrict__ partials, int32_t* __restrict__ bad) {
    float inter[KMAX], den[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) { inter[c] = 0.f; den[c] = 0.f; }
    float ce = 0.f;
    unsigned nvalid = 0;
    const int64_t total = (int64_t)B * HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t tg = __ldcs(target + i);
        if (tg == ignore) continue;                           // mask = target != ignore_index (:115)
        if (tg < 0 || tg >= K) {                              // neither a class nor the ignore label: counted (the reference raises)
            atomicAdd(&partials[2 * K + 2], 1.0);
            continue;
        }
        const int64_t b = i / HW, px = i - b * HW;
        const float* lp = logits + (b * K) * HW + px;
        float v[KMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            v[c] = (c < K) ? __ldcs(lp + (int64_t)c * HW) : -INFINITY;
            mx = fmaxf(mx, v[c]);
        }
        float z = 0.f, vt = 0.f;                              // vt = logit_t - max
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            vt = (c == (int)tg) ? v[c] - mx : vt;
            v[c] = (c < K) ? __expf(v[c] - mx) : 0.f;
            z += v[c];
        }
        const float inv = 1.0f / z;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            const float p = v[c] * inv;
            const bool is_t = (c == (int)tg);
            den[c] += p * p + (is_t ? 1.0f : 0.0f);
            inter[c] += is_t ? p : 0.0f;
        }
        ce += logf(z) - vt;                                   // log-sum-exp form: finite even when softmax_t underflows
        ++nvalid;
    }
    // block reduction (float64): every warp shuffles all 2K+2 values down, one barrier, 2K+2 threads finish
    __shared__ double s_red[8][2 * KMAX + 2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 2 * KMAX + 2; ++q) {
        double val = (q < KMAX) ? (double)inter[q < KMAX ? q : 0]
                   : (q < 2 * KMAX) ? (double)den[q < 2 * KMAX ? (q - KMAX >= 0 ? q - KMAX : 0) : 0]
                   : (q == 2 * KMAX) ? (double)ce : (double)nvalid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) s_red[w][q] = val;
    }
    __syncthreads();
    if (threadIdx.x < 2 * KMAX + 2) {
        const int q = threadIdx.x;
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += s_red[k][q];
        // slot layout of `partials`: inter[K], denom[K], ce_sum, n_valid
        int dst = -1;
        if (q < KMAX) { if (q < K) dst = q; }
        else if (q < 2 * KMAX) { if (q - KMAX < K) dst = K + (q - KMAX); }
        else dst = 2 * K + (q - 2 * KMAX);
        if (dst >= 0 && a != 0.0) atomicAdd(&partials[dst], a);
    }
}

__global__ void k_dice_ce_finish(const double* __restrict__ partials, int K, int64_t ignore, float w_dice, float w_ce,
                                 float* __restrict__ losses) {
    if (threadIdx.x || blockIdx.x) return;
    double dice = 0.0;
    for (int c = 0; c < K; ++c) {                            // BinaryDiceLoss, smooth = 1, p = 2 (:80-90)
        if ((int64_t)c == ignore) continue;                  // `if i != self.ignore_index` (:127) -- still divided by K below
        dice += 1.0 - (2.0 * partials[c] + 1.0) / (partials[K + c] + 1.0);
    }
    dice /= (double)K;                                       // total_loss / target.shape[1] (:135)
    const double ce = partials[2 * K] / partials[2 * K + 1]; // CrossEntropyLoss mean over valid (0/0 -> NaN)
    losses[0] = (float)dice;
    losses[1] = (float)ce;
    losses[2] = (float)((double)w_dice * dice + (double)w_ce * ce);
}

template <int KMAX>
__global__ void __launch_bounds__(256)
k_dice_ce_bwd(const float* __restrict__ logits, const int64_t* __restrict__ target, int B, int K, int64_t HW,
              int64_t ignore, const double* __restrict__ partials, float w_dice, float w_ce,
              const float* __restrict__ grad_scale, float* __restrict__ d_logits) {
    __shared__ float s_a[KMAX], s_b[KMAX];   // dL_dice/dp_c = s_a[c] * t_c + s_b[c] * p_c
    __shared__ float s_ce;
    if (threadIdx.x < KMAX) {
        const int c = threadIdx.x;
        if (c < K) {
            const double D1 = partials[K + c] + 1.0, N1 = 2.0 * partials[c] + 1.0;
            const bool skip = (int64_t)c == ignore;          // the Dice term of class `ignore_index` is not part of the loss
            s_a[c] = skip ? 0.f : (float)(-2.0 / D1 / (double)K);
            s_b[c] = skip ? 0.f : (float)(2.0 * N1 / (D1 * D1) / (double)K);
        } else {
            s_a[c] = 0.f; s_b[c] = 0.f;
        }
    }
    if (threadIdx.x == 0) s_ce = (float)((double)w_ce / partials[2 * K + 1]);
    __syncthreads();
    const float gs = grad_scale ? grad_scale[0] : 1.0f;
    const float ce_scale = s_ce;
    const int64_t total = (int64_t)B * HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t tg = __ldcs(target + i);
        const int64_t b = i / HW, px = i - b * HW;
        const float* lp = logits + (b * K) * HW + px;
        float* dp = d_logits + (b * K) * HW + px;
        if (tg == ignore || tg < 0 || tg >= K) {
            for (int c = 0; c < K; ++c) __stcs(dp + (int64_t)c * HW, 0.0f);
            continue;
        }
        float v[KMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            v[c] = (c < K) ? __ldcs(lp + (int64_t)c * HW) : -INFINITY;
            mx = fmaxf(mx, v[c]);
        }
        float z = 0.f;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) { v[c] = (c < K) ? __expf(v[c] - mx) : 0.f; z += v[c]; }
        const float inv = 1.0f / z;
        float dot = 0.f;
        float gp[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            v[c] *= inv;
            gp[c] = s_a[c] * ((c == (int)tg) ? 1.0f : 0.0f) + s_b[c] * v[c];
            dot += gp[c] * v[c];
        }
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            if (c < K) {
                const float t = (c == (int)tg) ? 1.0f : 0.0f;
                const float g = w_dice * v[c] * (gp[c] - dot) + ce_scale * (v[c] - t);
                __stcs(dp + (int64_t)c * HW, gs * g);
            }
        }
    }
}

// =============================================================================================
// a15 superpixel mean-pool.  feat NCHW is streamed once (4*B*Cf*H*W bytes): a warp owns 32*VEC
// consecutive pixels, keeps their (channel independent) segment ids in registers and loops over the
// channels of its channel block.  Neighbouring pixels mostly share a superpixel, so each channel value
// is reduced over runs of equal ids with a segmented warp shuffle and only run heads issue a
// red.global.add.f32 into pooled[id', c].  A second tiny kernel divides by (count + 1e-6).
// =============================================================================================
constexpr int kPoolWarps = 8;

template <int VEC>
struct PixelRun {
    long long id[VEC];   // id' or -1
    bool uniform;        // all VEC ids equal and valid
    bool head;           // first lane of a run of uniform lanes with equal id
    int end;             // last lane of my run
};

template <int VEC>
__device__ __forceinline__ PixelRun<VEC> load_runs(const int64_t* __restrict__ seg, int b, int64_t HW, int64_t p0,
                                                   int S, int64_t M, int32_t* status) {
    PixelRun<VEC> r;
    const int lane = threadIdx.x & 31;
    bool uni = true;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        long long id = -1;
        if (p0 + k < HW) {
            id = seg[(int64_t)b * HW + p0 + k] + (long long)b * S;      // pretrain_trainer.py:446-449
            if (id < 0 || id >= M) { if (status) *status = 1; id = -1; }
        }
        r.id[k] = id;
        uni = uni && (id >= 0) && (id == r.id[0]);
    }
    r.uniform = uni;
    const long long prev_id = __shfl_up_sync(0xffffffffu, r.id[0], 1);
    const bool prev_uni = __shfl_up_sync(0xffffffffu, (int)uni, 1) != 0;
    r.head = (lane == 0) || !uni || !prev_uni || (prev_id != r.id[0]);
    const unsigned hm = __ballot_sync(0xffffffffu, r.head);
    const unsigned above = (lane == 31) ? 0u : (hm & ~((2u << lane) - 1u));
    r.end = above ? (__ffs(above) - 2) : 31;
    return r;
}

__device__ __forceinline__ float seg_reduce(float val, int lane, int end) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, val, o);
        if (lane + o <= end) val += t;
    }
    return val;
}

template <int VEC>
__global__ void __launch_bounds__(kPoolWarps * 32)
k_segpool_sum(const float* __restrict__ feat, const int64_t* __restrict__ seg, int B, int Cf, int64_t HW, int S,
              int64_t M, int ch_per_block, float* __restrict__ pooled, float* __restrict__ counts,
              int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int64_t p0 = ((int64_t)blockIdx.x * kPoolWarps + w) * (32 * VEC) + (int64_t)lane * VEC;
    if (p0 - (int64_t)lane * VEC >= HW) return;   // whole warp out of range
    const PixelRun<VEC> r = load_runs<VEC>(seg, b, HW, p0, S, M, status);
    const int c0 = blockIdx.y * ch_per_block;
    const int c1 = min(Cf, c0 + ch_per_block);

    if (blockIdx.y == 0) {   // counts once per pixel tile
        float cnt = r.uniform ? (float)VEC : 0.f;
        if (!r.uniform) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) if (r.id[k] >= 0) atomicAdd(counts + r.id[k], 1.0f);
        }
        cnt = seg_reduce(cnt, lane, r.end);
        if (r.head && r.uniform) atomicAdd(counts + r.id[0], cnt);
    }
    const float* fp = feat + ((int64_t)b * Cf + c0) * HW + p0;
    for (int c = c0; c < c1; ++c, fp += HW) {
        float v[VEC];
        if (VEC == 4) {
            const float4 q = (p0 + 3 < HW) ? __ldcs(reinterpret_cast<const float4*>(fp)) : make_float4(0, 0, 0, 0);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[k] = (p0 + k < HW) ? __ldcs(fp + k) : 0.f;
        }
        float val = 0.f;
        if (r.uniform) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) val += v[k];
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) if (r.id[k] >= 0) atomicAdd(pooled + r.id[k] * Cf + c, v[k]);
        }
        val = seg_reduce(val, lane, r.end);
        if (r.head && r.uniform) atomicAdd(pooled + r.id[0] * Cf + c, val);
    }
}

__global__ void k_segpool_finish(float* __restrict__ pooled, const float* __restrict__ counts, int64_t M, int Cf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * Cf) return;
    pooled[i] = pooled[i] / (counts[i / Cf] + 1e-6f);        // k / (sum(one_hot, 1)[:, None] + 1e-6)
}

template <int VEC>
__global__ void __launch_bounds__(kPoolWarps * 32)
k_segpool_bwd(const float* __restrict__ d_pooled, const int64_t* __restrict__ seg, const float* __restrict__ counts,
              int B, int Cf, int64_t HW, int S, int64_t M, int ch_per_block, float* __restrict__ d_feat) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int64_t p0 = ((int64_t)blockIdx.x * kPoolWarps + w) * (32 * VEC) + (int64_t)lane * VEC;
    if (p0 >= HW) return;
    long long id[VEC];
    float inv[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        id[k] = -1; inv[k] = 0.f;
        if (p0 + k < HW) {
            const long long v = seg[(int64_t)b * HW + p0 + k] + (long long)b * S;
            if (v >= 0 && v < M) { id[k] = v; inv[k] = 1.0f / (counts[v] + 1e-6f); }
        }
    }
    const int c0 = blockIdx.y * ch_per_block;
    const int c1 = min(Cf, c0 + ch_per_block);
    float* dp = d_feat + ((int64_t)b * Cf + c0) * HW + p0;
    for (int c = c0; c < c1; ++c, dp += HW) {
        float v[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = (id[k] >= 0) ? __ldg(d_pooled + id[k] * Cf + c) * inv[k] : 0.f;
        if (VEC == 4 && p0 + 3 < HW) {
            __stcs(reinterpret_cast<float4*>(dp), make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) if (p0 + k < HW) __stcs(dp + k, v[k]);
        }
    }
}

}  // namespace oess

using namespace oess;

OESS_API int oess_confusion(const int64_t* pred, const int64_t* gt, int64_t n, int K, int64_t ignore_label,
                            int64_t* conf, int32_t* status, oess_stream_t stream) {
    if (n < 0 || K <= 0 || (int64_t)K * K >= (1ll << 31)) return OESS_E_ARG;
    if (n == 0) return OESS_OK;
    if (!pred || !gt || !conf) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) OESS_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    OESS_KERNEL("k_confusion", st, k_confusion<<<(unsigned)blocks, 256, 0, st>>>(pred, gt, n, K, ignore_label, (unsigned long long*)conf, status));
    return OESS_OK;
}

OESS_API int oess_argmax_confusion(const float* logits, const int64_t* gt, int B, int K, int H, int W, int64_t ignore_label,
                                   int64_t* conf, int32_t* status, oess_stream_t stream) {
    if (B <= 0 || K <= 0 || K > 64 || H <= 0 || W <= 0 || !logits || !gt || !conf) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) OESS_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    const int64_t n = (int64_t)B * H * W;
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    OESS_KERNEL("k_argmax_confusion", st, k_argmax_confusion<<<(unsigned)blocks, 256, 0, st>>>(
        logits, gt, B, K, (int64_t)H * W, ignore_label, (unsigned long long*)conf, status));
    return OESS_OK;
}

static inline unsigned loss_grid(int64_t total) {
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

OESS_API int oess_dice_ce_partials(const float* logits, const int64_t* target, int B, int K, int H, int W,
                                   int64_t ignore_index, double* partials, oess_stream_t stream) {
    if (B <= 0 || K <= 0 || K > 64 || H <= 0 || W <= 0 || !logits || !target || !partials) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    OESS_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * (2 * K + 3), st));
    const unsigned g = loss_grid((int64_t)B * HW);
    OESS_KERNEL("k_dice_ce_partials", st, if (K <= 8) k_dice_ce_partials<8><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, nullptr);
    else if (K <= 16) k_dice_ce_partials<16><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, nullptr);
    else if (K <= 32) k_dice_ce_partials<32><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, nullptr);
    else k_dice_ce_partials<64><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, nullptr));
    return OESS_OK;
}

OESS_API int oess_dice_ce_finish_ex(const double* partials, int K, int64_t ignore_index, float w_dice, float w_ce,
                                    float* losses, oess_stream_t stream) {
    if (!partials || !losses || K <= 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("k_dice_ce_finish", st, k_dice_ce_finish<<<1, 32, 0, st>>>(partials, K, ignore_index, w_dice, w_ce, losses));
    return OESS_OK;
}

OESS_API int oess_dice_ce_finish(const double* partials, int K, float w_dice, float w_ce, float* losses,
                                 oess_stream_t stream) {
    return oess_dice_ce_finish_ex(partials, K, -(1ll << 62), w_dice, w_ce, losses, stream);
}

OESS_API int oess_dice_ce_bwd(const float* logits, const int64_t* target, int B, int K, int H, int W,
                              int64_t ignore_index, const double* partials, float w_dice, float w_ce,
                              const float* grad_scale, float* d_logits, oess_stream_t stream) {
    if (B <= 0 || K <= 0 || K > 64 || H <= 0 || W <= 0 || !logits || !target || !partials || !d_logits)
        return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    const unsigned g = loss_grid((int64_t)B * HW);
    OESS_KERNEL("k_dice_ce_bwd", st, if (K <= 8) k_dice_ce_bwd<8><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, w_dice, w_ce, grad_scale, d_logits);
    else if (K <= 16) k_dice_ce_bwd<16><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, w_dice, w_ce, grad_scale, d_logits);
    else if (K <= 32) k_dice_ce_bwd<32><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, w_dice, w_ce, grad_scale, d_logits);
    else k_dice_ce_bwd<64><<<g, 256, 0, st>>>(logits, target, B, K, HW, ignore_index, partials, w_dice, w_ce, grad_scale, d_logits));
    return OESS_OK;
}

OESS_API int oess_segpool_ws_bytes(int B, int Cf, int H, int W, int64_t M, size_t* ws_bytes) {
    if (!ws_bytes || B <= 0 || Cf <= 0 || H <= 0 || W <= 0 || M <= 0) return OESS_E_ARG;
    *ws_bytes = 0;   // accumulates straight into `pooled` / `counts`
    return OESS_OK;
}

static void pool_grid(int B, int Cf, int64_t HW, int vec, dim3* grid, int* ch_per_block) {
    const int64_t px_blocks = (HW + (int64_t)kPoolWarps * 32 * vec - 1) / ((int64_t)kPoolWarps * 32 * vec);
    // enough CTAs for >= 4 waves, but keep channel blocks long so the ids are reused
    int cb = Cf;
    while (cb > 8 && px_blocks * B * ((Cf + cb - 1) / cb) < (int64_t)kNumSMs * 8) cb = (cb + 1) / 2;
    *ch_per_block = cb;
    *grid = dim3((unsigned)px_blocks, (unsigned)((Cf + cb - 1) / cb), (unsigned)B);
}

OESS_API int oess_segpool_fwd(const float* feat, const int64_t* seg, int B, int Cf, int H, int W, int S, int64_t M,
                              float* pooled, float* counts, int32_t* status, void* ws, size_t ws_bytes,
                              oess_stream_t stream) {
    (void)ws; (void)ws_bytes;
    if (B <= 0 || Cf <= 0 || H <= 0 || W <= 0 || M <= 0 || B > 65535) return OESS_E_ARG;
    if (!feat || !seg || !pooled || !counts) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    OESS_CUDA(cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)M * Cf, st));
    OESS_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * (size_t)M, st));
    if (status) OESS_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    const bool vec4 = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
    dim3 grid; int cb;
    pool_grid(B, Cf, HW, vec4 ? 4 : 1, &grid, &cb);
    OESS_KERNEL("k_segpool_sum", st, if (vec4) k_segpool_sum<4><<<grid, kPoolWarps * 32, 0, st>>>(feat, seg, B, Cf, HW, S, M, cb, pooled, counts, status);
    else k_segpool_sum<1><<<grid, kPoolWarps * 32, 0, st>>>(feat, seg, B, Cf, HW, S, M, cb, pooled, counts, status));
    OESS_KERNEL("k_segpool_finish", st, k_segpool_finish<<<(unsigned)((M * Cf + 255) / 256), 256, 0, st>>>(pooled, counts, M, Cf));
    return OESS_OK;
}

OESS_API int oess_segpool_bwd(const float* d_pooled, const int64_t* seg, const float* counts, int B, int Cf, int H,
                              int W, int S, int64_t M, float* d_feat, oess_stream_t stream) {
    if (B <= 0 || Cf <= 0 || H <= 0 || W <= 0 || M <= 0 || B > 65535) return OESS_E_ARG;
    if (!d_pooled || !seg || !counts || !d_feat) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    const bool vec4 = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_feat) & 15) == 0);
    dim3 grid; int cb;
    pool_grid(B, Cf, HW, vec4 ? 4 : 1, &grid, &cb);
    OESS_KERNEL("k_segpool_bwd", st, if (vec4) k_segpool_bwd<4><<<grid, kPoolWarps * 32, 0, st>>>(d_pooled, seg, counts, B, Cf, HW, S, M, cb, d_feat);
    else k_segpool_bwd<1><<<grid, kPoolWarps * 32, 0, st>>>(d_pooled, seg, counts, B, Cf, HW, S, M, cb, d_feat));
    return OESS_OK;
}
