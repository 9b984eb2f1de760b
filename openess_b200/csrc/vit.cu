// vit.cu -- the non-GEMM kernels of the MaskCLIP ViT-B/16 forward (models/maskclip_model.py):
// patch extraction, token assembly, LayerNorm, multi-head attention, the head's L2 normalisation and the final bilinear
// resize of the logits.  The linears (patch embedding, in_proj, out_proj, FFN, head proj, text-embedding classifier) run
// on the tcgen05 GEMM (tc_gemm.cu, oess_gemm_tf32_ex with the GELU / residual epilogue).
//
// Attention: `oess_mha_fwd` below is the exact-fp32 variant on the FMA pipes (flash-style: 64-query x 64-key tiles, online
// softmax, no T x T matrix in memory; 1.0 ms per ViT-B/16 layer at B = 8, T = 1121); the default path of the model is the
// tcgen05 kernel in tc_mha.cu (`oess_mha_fwd_tc`, 0.13 ms), this one stays as the reference-precision variant the tests
// compare it with.
#include <math.h>

#include "common.cuh"

namespace oess {
namespace vit {

// ------------------------------------------------------------------------------------------------- patchify
__global__ void __launch_bounds__(256)
k_patchify(const float* __restrict__ img, int B, int C, int H, int W, int P, int h, int w, float* __restrict__ rows) {
    const int64_t KK = (int64_t)C * P * P;
    const int64_t total = (int64_t)B * h * w * KK;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / KK;
        int col = (int)(i - row * KK);
        const int kx = col % P; col /= P;
        const int ky = col % P;
        const int c = col / P;
        const int px = (int)(row % w);
        const int py = (int)((row / w) % h);
        const int b = (int)(row / ((int64_t)w * h));
        const int y = py * P + ky, x = px * P + kx;
        rows[i] = (y < H && x < W) ? img[(((int64_t)b * C + c) * H + y) * W + x] : 0.0f;   // 'corner' zero padding
    }
}

// ------------------------------------------------------------------------------------------------- cls + pos
__global__ void __launch_bounds__(256)
k_assemble(const float4* __restrict__ tok, const float4* __restrict__ cls, const float4* __restrict__ pos, int B, int T,
           int D4, float4* __restrict__ x) {
    const int64_t total = (int64_t)B * T * D4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D4);
        const int t = (int)((i / D4) % T);
        const int b = (int)(i / ((int64_t)D4 * T));
        const float4 a = t == 0 ? cls[d] : tok[((int64_t)b * (T - 1) + (t - 1)) * D4 + d];
        const float4 p = pos[(int64_t)t * D4 + d];
        x[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
}

// ------------------------------------------------------------------------------------------------- row kernels
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kMaxV4 = 8;   // D <= 1024: a lane keeps its D / 32 values in registers

// One warp per row; the row is read once (two-pass mean / variance out of registers, as torch's LayerNorm kernel).
__global__ void __launch_bounds__(256)
k_layernorm_rows(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 int64_t rows, int D, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nv = D >> 7;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[kMaxV4];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i)
        if (i < nv) {
            v[i] = xr[i * 32 + lane];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i)
        if (i < nv) {
            const float4 g = g4[i * 32 + lane], b = b4[i * 32 + lane];
            yr[i * 32 + lane] = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                            (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        }
}

__global__ void __launch_bounds__(256)
k_l2norm_rows(float* __restrict__ x, int64_t rows, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nv = D >> 7;
    float4* xr = reinterpret_cast<float4*>(x + row * D);
    float4 v[kMaxV4];
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i)
        if (i < nv) {
            v[i] = xr[i * 32 + lane];
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    const float n = sqrtf(warp_sum(q));          // feat / feat.norm(dim=1): no eps (maskclip_model.py:219)
#pragma unroll
    for (int i = 0; i < kMaxV4; ++i)
        if (i < nv) xr[i * 32 + lane] = make_float4(v[i].x / n, v[i].y / n, v[i].z / n, v[i].w / n);
}

// ------------------------------------------------------------------------------------------------- attention
constexpr int kHd = 64;          // head dim
constexpr int kTq = 64;          // queries per CTA
constexpr int kTk = 64;          // keys per tile
constexpr int kPs = 68;          // row stride of the probability tile (floats): half-warps hit different banks
constexpr int kMhaSmem = (kHd * kTq + kHd * kTk + kTk * kHd + kTq * kPs) * 4;

// grid (ceil(T / 64), heads, B), 256 threads: thread (ty, tx) = (tid / 16, tid % 16) owns a 4 x 4 patch of the 64 x 64
// score tile and of the 64 x 64 output tile.  Q and K tiles are stored [d][row] so the inner product reads two float4.
__global__ void __launch_bounds__(256)
k_mha_fwd(const float* __restrict__ qkv, int T, int heads, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                   // [64 d][64 q]
    float* Ks = Qs + kHd * kTq;         // [64 d][64 k]
    float* Vs = Ks + kHd * kTk;         // [64 k][64 d]
    float* Ps = Vs + kTk * kHd;         // [64 q][68]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * kTq;
    const int hh = blockIdx.y, b = blockIdx.z;
    const int D = heads * kHd;
    const int64_t rs = 3 * (int64_t)D;                                   // row stride of qkv
    const float* base = qkv + (int64_t)b * T * rs + hh * kHd;

    // Q tile, transposed: lanes = consecutive rows (conflict-free shared stores)
    {
        const int r = tid & 63;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int d4 = ((tid >> 6) + p * 4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q0 + r < T) v = *reinterpret_cast<const float4*>(base + (int64_t)(q0 + r) * rs + d4);
            Qs[(d4 + 0) * kTq + r] = v.x * 0.125f;                        // q * head_dim^-0.5 (exact: power of two)
            Qs[(d4 + 1) * kTq + r] = v.y * 0.125f;
            Qs[(d4 + 2) * kTq + r] = v.z * 0.125f;
            Qs[(d4 + 3) * kTq + r] = v.w * 0.125f;
        }
    }
    float o[4][4];
    float m[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.0f;
    }

    for (int k0 = 0; k0 < T; k0 += kTk) {
        {   // K tile transposed, V tile as is
            const int r = tid & 63;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int d4 = ((tid >> 6) + p * 4) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + r < T) v = *reinterpret_cast<const float4*>(base + D + (int64_t)(k0 + r) * rs + d4);
                Ks[(d4 + 0) * kTk + r] = v.x;
                Ks[(d4 + 1) * kTk + r] = v.y;
                Ks[(d4 + 2) * kTk + r] = v.z;
                Ks[(d4 + 3) * kTk + r] = v.w;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int idx = tid + p * 256;                 // 1024 float4 of the V tile
                const int rv = idx >> 4, c4 = (idx & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + rv < T) v = *reinterpret_cast<const float4*>(base + 2 * D + (int64_t)(k0 + rv) * rs + c4);
                *reinterpret_cast<float4*>(Vs + rv * kHd + c4) = v;
            }
        }
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll 8
        for (int d = 0; d < kHd; ++d) {
            const float4 qa = *reinterpret_cast<const float4*>(Qs + d * kTq + 4 * ty);
            const float4 ka = *reinterpret_cast<const float4*>(Ks + d * kTk + 4 * tx);
            const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {ka.x, ka.y, ka.z, ka.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + 4 * tx + j >= T) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i][j] = -INFINITY;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
            for (int sh = 8; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
            const float mn = fmaxf(m[i], mx);                  // finite: every tile has at least one valid key
            const float alpha = __expf(m[i] - mn);             // first tile: exp(-inf) = 0
            m[i] = mn;
            float4 p;
            p.x = __expf(s[i][0] - mn);
            p.y = __expf(s[i][1] - mn);
            p.z = __expf(s[i][2] - mn);
            p.w = __expf(s[i][3] - mn);
            l[i] = l[i] * alpha + ((p.x + p.y) + (p.z + p.w));
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
            *reinterpret_cast<float4*>(Ps + (4 * ty + i) * kPs + 4 * tx) = p;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kTk; k += 4) {
            float pv[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t4 = *reinterpret_cast<const float4*>(Ps + (4 * ty + i) * kPs + k);
                pv[i][0] = t4.x; pv[i][1] = t4.y; pv[i][2] = t4.z; pv[i][3] = t4.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 va = *reinterpret_cast<const float4*>(Vs + (k + kk) * kHd + 4 * tx);
                const float vv[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i][kk], vv[j], o[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float ls = l[i];
#pragma unroll
        for (int sh = 8; sh > 0; sh >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, sh);
        const int q = q0 + 4 * ty + i;
        if (q < T) {
            const float inv = 1.0f / ls;
            *reinterpret_cast<float4*>(out + ((int64_t)b * T + q) * D + hh * kHd + 4 * tx) =
                make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        }
    }
}

// ------------------------------------------------------------------------------------------------- logits resize
// F.interpolate(mode='bilinear', align_corners=False) of a channels-last [B, h, w, K] map to [B, K, H, W].
__global__ void __launch_bounds__(256)
k_bilinear_tokens_to_nchw(const float* __restrict__ tok, int B, int h, int w, int K, int H, int W, float* __restrict__ out) {
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const int64_t total = (int64_t)B * K * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int k = (int)((i / ((int64_t)W * H)) % K);
        const int b = (int)(i / ((int64_t)W * H * K));
        const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.0f);
        const float fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.0f);
        const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
        const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float* t = tok + (int64_t)b * h * w * K + k;
        const float v00 = t[((int64_t)y0 * w + x0) * K], v01 = t[((int64_t)y0 * w + x1) * K];
        const float v10 = t[((int64_t)y1 * w + x0) * K], v11 = t[((int64_t)y1 * w + x1) * K];
        out[i] = (1.0f - ly) * ((1.0f - lx) * v00 + lx * v01) + ly * ((1.0f - lx) * v10 + lx * v11);
    }
}

static inline unsigned grid_for(int64_t total, int threads) {
    const int64_t g = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vit
}  // namespace oess

using namespace oess;

OESS_API int oess_vit_patchify(const float* img, int B, int C, int H, int W, int P, float* rows, oess_stream_t stream) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || P <= 0) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!img || !rows) return OESS_E_ARG;
    const int h = (H + P - 1) / P, w = (W + P - 1) / P;
    const int64_t total = (int64_t)B * h * w * C * P * P;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("vit_patchify", st, vit::k_patchify<<<vit::grid_for(total, 256), 256, 0, st>>>(img, B, C, H, W, P, h, w, rows));
    return OESS_OK;
}

OESS_API int oess_vit_assemble(const float* tok, const float* cls, const float* pos, int B, int T, int D, float* x,
                               oess_stream_t stream) {
    if (B < 0 || T < 1 || D <= 0 || (D & 3)) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!cls || !pos || !x || (T > 1 && !tok)) return OESS_E_ARG;
    if (((uintptr_t)tok | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)x) & 15) return OESS_E_ARG;
    const int64_t total = (int64_t)B * T * (D / 4);
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("vit_assemble", st, vit::k_assemble<<<vit::grid_for(total, 256), 256, 0, st>>>(
        (const float4*)tok, (const float4*)cls, (const float4*)pos, B, T, D / 4, (float4*)x));
    return OESS_OK;
}

OESS_API int oess_layernorm_rows(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int D,
                                 float* y, oess_stream_t stream) {
    if (rows < 0 || D <= 0 || (D & 127) || D > 128 * vit::kMaxV4) return OESS_E_ARG;
    if (rows == 0) return OESS_OK;
    if (!x || !gamma || !beta || !y) return OESS_E_ARG;
    if (((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)y) & 15) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("layernorm_rows", st, vit::k_layernorm_rows<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, gamma, beta, eps, rows, D, y));
    return OESS_OK;
}

OESS_API int oess_l2norm_rows(float* x, int64_t rows, int D, oess_stream_t stream) {
    if (rows < 0 || D <= 0 || (D & 127) || D > 128 * vit::kMaxV4) return OESS_E_ARG;
    if (rows == 0) return OESS_OK;
    if (!x || ((uintptr_t)x & 15)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("l2norm_rows", st, vit::k_l2norm_rows<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, D));
    return OESS_OK;
}

OESS_API int oess_mha_fwd(const float* qkv, int B, int T, int heads, float* out, oess_stream_t stream) {
    if (B < 0 || T <= 0 || heads <= 0 || heads > 65535 || B > 65535) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!qkv || !out || (((uintptr_t)qkv | (uintptr_t)out) & 15)) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaFuncSetAttribute(vit::k_mha_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, vit::kMhaSmem));
    const dim3 grid((unsigned)((T + vit::kTq - 1) / vit::kTq), (unsigned)heads, (unsigned)B);
    OESS_KERNEL("mha_fwd", st, vit::k_mha_fwd<<<grid, 256, vit::kMhaSmem, st>>>(qkv, T, heads, out));
    return OESS_OK;
}

OESS_API int oess_bilinear_tokens_to_nchw(const float* tok, int B, int h, int w, int K, int H, int W, float* out,
                                          oess_stream_t stream) {
    if (B < 0 || h <= 0 || w <= 0 || K <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!tok || !out) return OESS_E_ARG;
    const int64_t total = (int64_t)B * K * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_KERNEL("bilinear_tokens_to_nchw", st, vit::k_bilinear_tokens_to_nchw<<<vit::grid_for(total, 256), 256, 0, st>>>(
        tok, B, h, w, K, H, W, out));
    return OESS_OK;
}
