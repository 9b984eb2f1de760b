// infonce.cu -- a16: utils/loss_functions.py:138-153 NCELoss (PointInfoNCE).
//   s_ij = (k_i . q_j) / T;   loss = mean_i (logsumexp_j s_ij - s_ii);   optional gradients dk, dq.
// fp32 SIMT tiles (the reference computes in fp32; M = #superpixels in the batch <= a few thousand, D = 256:
// a few GFLOP per step, well below the dense encoders).  One tiled kernel, three modes:
//   LSE : A = k (64 rows resident in smem), B = q tiles -> per-row (max, sum-exp) partials + the diagonal
//   DK  : A = k, B = q:  dk_a += sum_b G_ab q_b,  G_ab = (exp(s_ab - lse_a) - [a==b]) / (M T)
//   DQ  : A = q, B = k:  dq_a += sum_b G_ba k_b,  G_ba = (exp(s_ba - lse_b) - [a==b]) / (M T)
// Parallelism: grid = (row tiles) x (column splits); a CTA handles column tiles s, s+nsplit, ... and writes
// PARTIAL results that small combine kernels reduce in a fixed order (deterministic, no float atomics).
// M = 800 (batch 8 x 100 superpixels) would otherwise run on 13 of 148 SMs.
#include <stdlib.h>

#include "common.cuh"

namespace oess {

constexpr int kNT = 64;  // tile of rows / cols
enum { NCE_LSE = 0, NCE_DK = 1, NCE_DQ = 2 };

template <int MODE, int DC>   // feature width padded to D = 16 * DC columns
__global__ void __launch_bounds__(256)
k_infonce(const float* __restrict__ A, const float* __restrict__ Bm, int64_t M, int Dr, float inv_T,
          const float* __restrict__ lse, float* __restrict__ pm, float* __restrict__ pl, float* __restrict__ diag,
          float* __restrict__ part) {
    constexpr int D = 16 * DC;
    constexpr int LD = D + 1;                 // padded rows: conflict-free column access
    extern __shared__ float smem[];
    float* As = smem;                         // [64][LD]
    float* Bs = As + kNT * LD;                // [64][LD]
    float* Gs = Bs + kNT * LD;                // [64][65]
    __shared__ float s_m[kNT], s_l[kNT];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t a0 = (int64_t)blockIdx.x * kNT;
    const int split = blockIdx.y, nsplit = gridDim.y;
    for (int i = tid; i < kNT * D; i += 256) {
        const int r = i / D, d = i - r * D;
        As[r * LD + d] = (a0 + r < M && d < Dr) ? A[(a0 + r) * Dr + d] : 0.f;   // zero padded to D
    }
    if (tid < kNT) { s_m[tid] = -INFINITY; s_l[tid] = 0.f; }
    float acc[4][DC];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < DC; ++c) acc[r][c] = 0.f;
    const float inv_MT = inv_T / (float)M;

    for (int64_t b0 = (int64_t)split * kNT; b0 < M; b0 += (int64_t)nsplit * kNT) {
        __syncthreads();
        for (int i = tid; i < kNT * D; i += 256) {
            const int r = i / D, d = i - r * D;
            Bs[r * LD + d] = (b0 + r < M && d < Dr) ? Bm[(b0 + r) * Dr + d] : 0.f;
        }
        __syncthreads();
        // S tile: rows ty*4+r, cols tx+16*c
        float s[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            float av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) av[r] = As[(ty * 4 + r) * LD + d];
#pragma unroll
            for (int c = 0; c < 4; ++c) bv[c] = Bs[(tx + 16 * c) * LD + d];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[r][c] = fmaf(av[r], bv[c], s[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ra = ty * 4 + r, cb = tx + 16 * c;
                const int64_t ga = a0 + ra, gb = b0 + cb;
                float v = s[r][c] * inv_T;
                if (MODE == NCE_LSE) {
                    if (gb >= M) v = -INFINITY;
                } else {
                    // row index of the logits matrix is the k index: a for DK, b for DQ
                    const float l = (MODE == NCE_DK) ? ((ga < M) ? lse[ga] : 0.f) : ((gb < M) ? lse[gb] : 0.f);
                    v = (ga < M && gb < M) ? (__expf(v - l) - (ga == gb ? 1.0f : 0.0f)) * inv_MT : 0.f;
                }
                Gs[ra * 65 + cb] = v;
            }
        __syncthreads();
        if (MODE == NCE_LSE) {
            // 4 threads per row, 16 columns each: online (max, sum-exp)
            const int row = tid >> 2, part4 = tid & 3;
            float mx = -INFINITY;
            for (int c = part4 * 16; c < part4 * 16 + 16; ++c) mx = fmaxf(mx, Gs[row * 65 + c]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_old = s_m[row];
            const float m_new = fmaxf(m_old, mx);
            float sum = 0.f;
            for (int c = part4 * 16; c < part4 * 16 + 16; ++c) sum += __expf(Gs[row * 65 + c] - m_new);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            __syncwarp();                      // all 4 lanes of the row have read s_m[row] before lane 0 rewrites it
            if (part4 == 0) {
                s_l[row] = s_l[row] * __expf(m_old - m_new) + sum;
                s_m[row] = m_new;
                const int64_t ga = a0 + row;
                if (ga < M && ga >= b0 && ga < b0 + kNT) diag[ga] = Gs[row * 65 + (int)(ga - b0)];
            }
        } else {
            // acc[r][c] += sum_b G[ra][b] * B[b][tx + 16 c]
#pragma unroll 2
            for (int bb = 0; bb < kNT; ++bb) {
                float gv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) gv[r] = Gs[(ty * 4 + r) * 65 + bb];
#pragma unroll
                for (int c = 0; c < DC; ++c) {
                    const float bval = Bs[bb * LD + tx + 16 * c];
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r][c] = fmaf(gv[r], bval, acc[r][c]);
                }
            }
        }
    }
    __syncthreads();
    if (MODE == NCE_LSE) {
        if (tid < kNT && a0 + tid < M) {
            pm[(int64_t)split * M + a0 + tid] = s_m[tid];
            pl[(int64_t)split * M + a0 + tid] = s_l[tid];
        }
    } else {
        float* p = part + (int64_t)split * M * Dr;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t ga = a0 + ty * 4 + r;
            if (ga >= M) continue;
#pragma unroll
            for (int c = 0; c < DC; ++c)
                if (tx + 16 * c < Dr) p[ga * Dr + tx + 16 * c] = acc[r][c];
        }
    }
}

// lse_i from the per-split (max, sum-exp) partials, fixed split order
__global__ void k_infonce_lse(const float* __restrict__ pm, const float* __restrict__ pl, int64_t M, int nsplit,
                              float* __restrict__ lse) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, pm[(int64_t)s * M + i]);
    float l = 0.f;
    for (int s = 0; s < nsplit; ++s) l += pl[(int64_t)s * M + i] * __expf(pm[(int64_t)s * M + i] - m);
    lse[i] = m + __logf(l);
}

__global__ void __launch_bounds__(1024)
k_infonce_loss(const float* __restrict__ lse, const float* __restrict__ diag, int64_t M, float* __restrict__ loss) {
    __shared__ double s_red[32];
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < M; i += 1024) a += (double)lse[i] - (double)diag[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += s_red[k];
        loss[0] = (float)(t / (double)M);                    // CrossEntropyLoss mean reduction
    }
}

// out[i] = sum_s part[s][i], fixed order
__global__ void k_infonce_reduce(const float* __restrict__ part, int64_t n, int nsplit, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int s = 0; s < nsplit; ++s) a += part[(int64_t)s * n + i];
    out[i] = a;
}

static int nce_splits(int64_t M) {
    const int64_t tiles = (M + kNT - 1) / kNT;
    int64_t s = (2 * kNumSMs + tiles - 1) / tiles;   // aim at >= 2 CTAs per SM
    if (s > tiles) s = tiles;
    if (s < 1) s = 1;
    return (int)s;
}

struct NceWs {
    float *lse, *diag, *pm, *pl, *part;
    size_t bytes;
};
static NceWs nce_carve(void* ws, int64_t M, int D) {
    NceWs r{};
    WsCarver c(ws);
    const int ns = nce_splits(M);
    r.lse = c.take<float>((size_t)M);
    r.diag = c.take<float>((size_t)M);
    r.pm = c.take<float>((size_t)M * ns);
    r.pl = c.take<float>((size_t)M * ns);
    r.part = c.take<float>((size_t)M * ns * D);
    r.bytes = c.total();
    return r;
}

template <int DC>
static int run_infonce(const float* k, const float* q, int64_t M, int Dr, float T, float* loss, float* dk, float* dq,
                       const NceWs& w, cudaStream_t st) {
    constexpr int D = 16 * DC;
    const size_t smem = sizeof(float) * (2 * kNT * (D + 1) + kNT * 65);
    const int ns = nce_splits(M);
    const dim3 grid((unsigned)((M + kNT - 1) / kNT), (unsigned)ns);
    const unsigned eb = (unsigned)((M * Dr + 255) / 256);
    OESS_CUDA(cudaFuncSetAttribute(k_infonce<NCE_LSE, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OESS_KERNEL("k_infonce_lse_tiles", st, k_infonce<NCE_LSE, DC><<<grid, 256, smem, st>>>(
        k, q, M, Dr, 1.0f / T, nullptr, w.pm, w.pl, w.diag, nullptr));
    OESS_KERNEL("k_infonce_lse", st, k_infonce_lse<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(w.pm, w.pl, M, ns, w.lse));
    OESS_KERNEL("k_infonce_loss", st, k_infonce_loss<<<1, 1024, 0, st>>>(w.lse, w.diag, M, loss));
    if (dk) {
        OESS_CUDA(cudaFuncSetAttribute(k_infonce<NCE_DK, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OESS_KERNEL("k_infonce_dk_tiles", st, k_infonce<NCE_DK, DC><<<grid, 256, smem, st>>>(
            k, q, M, Dr, 1.0f / T, w.lse, nullptr, nullptr, nullptr, w.part));
        OESS_KERNEL("k_infonce_reduce", st, k_infonce_reduce<<<eb, 256, 0, st>>>(w.part, M * Dr, ns, dk));
    }
    if (dq) {
        OESS_CUDA(cudaFuncSetAttribute(k_infonce<NCE_DQ, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OESS_KERNEL("k_infonce_dq_tiles", st, k_infonce<NCE_DQ, DC><<<grid, 256, smem, st>>>(
            q, k, M, Dr, 1.0f / T, w.lse, nullptr, nullptr, nullptr, w.part));
        OESS_KERNEL("k_infonce_reduce", st, k_infonce_reduce<<<eb, 256, 0, st>>>(w.part, M * Dr, ns, dq));
    }
    return OESS_OK;
}


// ---------------------------------------------------------------------------------------------
// Tensor-core formulation for large M (BASELINE config 3: M = 3 200 superpixels of an exact-global batch of 32): the three
// contractions  S = k q^T,  dk = G q,  dq = G^T k  (2 M^2 D flops each) run on the tcgen05 GEMM (oess_gemm_tf32) instead of
// fp32 FMA tiles.  The reference computes in fp32, and the logits are divided by T = 0.07, so plain TF32 (10-bit
// mantissa) would cost ~3 digits of the loss: every operand is split into TF32-exact parts  a = hi + lo
// (hi = rna_tf32(a), lo = rna_tf32(a - hi))  and the product is  hi_a hi_b + hi_a lo_b + lo_a hi_b  -- ONE GEMM over a
// 3x longer K with the parts concatenated ("3xTF32": the dropped lo_a lo_b term is 2^-22 relative, accumulation is fp32).
//   1. k3 = [hi hi lo](k), q3 = [hi lo hi](q)            [Mp, 3 D]  (Mp = M rounded up to 4, padded rows are zero)
//   2. S  = k3 q3^T                                        [M, Mp]     tcgen05
//   3. lse_i, loss (one CTA per row, row in registers, deterministic final reduction in double)
//   4. G3 = [hi hi lo](G),  G_ij = (exp(S_ij / T - lse_i) - [i == j]) / (M T)      [M, 3 Mp]
//   5. dk = G3 qT3^T   with qT3 = [hi lo hi](q^T)  [D, 3 Mp]                           tcgen05
//   6. S' = q3' k3'^T = S^T  (second small GEMM: cheaper and simpler than transposing G), G3 <- [hi hi lo](G^T) from S' with
//      lse indexed by column,  dq = G3 kT3^T                                             tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// out[i][0:D | D:2D | 2D:3D] = pattern 0: (hi, hi, lo), pattern 1: (hi, lo, hi) of a[i][:], rows >= M zero.  a: [M, D]
__global__ void k_nce_split3(const float* __restrict__ a, int64_t M, int D, int64_t Mp, int pattern, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Mp * D) return;
    const int64_t i = idx / D;
    const int d = (int)(idx - i * D);
    const float v = i < M ? a[idx] : 0.0f;
    const float hi = rna_tf32(v), lo = rna_tf32(v - hi);
    float* o = out + i * 3 * D;
    o[d] = hi;
    o[D + d] = pattern ? lo : hi;
    o[2 * D + d] = pattern ? hi : lo;
}

// out[d][0:Mp | Mp:2Mp | 2Mp:3Mp] = (hi, lo, hi) of a[:, d] (a: [M, D]; columns >= M zero).  32 x 32 shared-memory transpose.
__global__ void __launch_bounds__(256)
k_nce_split3_t(const float* __restrict__ a, int64_t M, int D, int64_t Mp, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = i0 + r;
        const int d = d0 + tx;
        tile[r][tx] = (i < M && d < D) ? a[i * D + d] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r;
        const int64_t i = i0 + tx;
        if (d < D && i < Mp) {
            const float v = tile[tx][r];
            const float hi = rna_tf32(v), lo = rna_tf32(v - hi);
            float* o = out + (int64_t)d * 3 * Mp;
            o[i] = hi;
            o[Mp + i] = lo;
            o[2 * Mp + i] = hi;
        }
    }
}

// one CTA per row i of S [M, ld]: lse_i = logsumexp_j<M (S_ij * inv_T), rowloss_i = lse_i - S_ii * inv_T
__global__ void __launch_bounds__(256)
k_nce_rows(const float* __restrict__ S, int64_t M, int64_t ld, float inv_T, float* __restrict__ lse, float* __restrict__ rowloss) {
    __shared__ float s_red[8];
    const int64_t i = blockIdx.x;
    const float* row = S + i * ld;
    float m = -INFINITY;
    for (int64_t j = threadIdx.x; j < M; j += 256) m = fmaxf(m, row[j] * inv_T);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    float l = 0.0f;
    for (int64_t j = threadIdx.x; j < M; j += 256) l += expf(row[j] * inv_T - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        const float lg = logf(t);
        lse[i] = m + lg;
        rowloss[i] = (m - row[i] * inv_T) + lg;      // lse - s_ii without cancelling two numbers of magnitude 1 / T
    }
}

__global__ void __launch_bounds__(1024)
k_nce_loss(const float* __restrict__ rowloss, int64_t M, float* __restrict__ loss) {
    __shared__ double s_red[32];
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < M; i += 1024) a += (double)rowloss[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += s_red[k];
        loss[0] = (float)(t / (double)M);                    // CrossEntropyLoss mean reduction
    }
}

// G3[r][0:Mp | Mp:2Mp | 2Mp:3Mp] = (hi, hi, lo) of G[r][c] = (exp(X[r][c] * inv_T - lse[by_col ? c : r]) - [r == c]) * inv_MT,
// columns >= M zero.  X: [M, ld]
__global__ void __launch_bounds__(256)
k_nce_grad_split(const float* __restrict__ X, int64_t M, int64_t Mp, int64_t ld, float inv_T, float inv_MT,
                 const float* __restrict__ lse, int by_col, float* __restrict__ G3) {
    const int64_t r = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= Mp) return;
    float g = 0.0f;
    if (c < M) g = (expf(X[r * ld + c] * inv_T - lse[by_col ? c : r]) - (r == c ? 1.0f : 0.0f)) * inv_MT;
    const float hi = rna_tf32(g), lo = rna_tf32(g - hi);
    float* o = G3 + r * 3 * Mp;
    o[c] = hi;
    o[Mp + c] = hi;
    o[2 * Mp + c] = lo;
}

struct NceTcWs {
    float *k3, *q3, *qa3, *kb3, *qT3, *kT3, *S, *G3, *lse, *rowloss;
    size_t bytes;
};
static NceTcWs nce_tc_carve(void* ws, int64_t M, int D) {
    NceTcWs r{};
    WsCarver c(ws);
    const int64_t Mp = (M + 3) & ~(int64_t)3;
    r.k3 = c.take<float>((size_t)Mp * 3 * D);     // (hi hi lo)(k): A operand of S
    r.q3 = c.take<float>((size_t)Mp * 3 * D);     // (hi lo hi)(q): B operand of S
    r.qa3 = c.take<float>((size_t)Mp * 3 * D);    // (hi hi lo)(q): A operand of S^T
    r.kb3 = c.take<float>((size_t)Mp * 3 * D);    // (hi lo hi)(k): B operand of S^T
    r.qT3 = c.take<float>((size_t)D * 3 * Mp);
    r.kT3 = c.take<float>((size_t)D * 3 * Mp);
    r.S = c.take<float>((size_t)M * Mp);
    r.G3 = c.take<float>((size_t)M * 3 * Mp);
    r.lse = c.take<float>((size_t)M);
    r.rowloss = c.take<float>((size_t)M);
    r.bytes = c.total();
    return r;
}

static bool nce_use_tc(int64_t M, int D) {
    static const int mode = [] { const char* e = getenv("OESS_INFONCE"); return e ? (e[0] == 's' ? 0 : e[0] == 't' ? 2 : 1) : 1; }();
    if (mode == 0 || (D & 3) || M >= (1ll << 30)) return false;
    return mode == 2 ? true : M >= 1024;      // below ~1 k rows the fp32 tiles win (launch count, operand splitting)
}

static int run_infonce_tc(const float* k, const float* q, int64_t M, int D, float T, float* loss, float* dk, float* dq,
                          const NceTcWs& w, cudaStream_t st) {
    const int64_t Mp = (M + 3) & ~(int64_t)3;
    const float inv_T = 1.0f / T, inv_MT = inv_T / (float)M;
    const unsigned sb = (unsigned)((Mp * D + 255) / 256);
    OESS_KERNEL("nce_split3", st, k_nce_split3<<<sb, 256, 0, st>>>(k, M, D, Mp, 0, w.k3));
    OESS_KERNEL("nce_split3", st, k_nce_split3<<<sb, 256, 0, st>>>(q, M, D, Mp, 1, w.q3));
    int rc = oess_gemm_tf32(w.k3, w.q3, nullptr, w.S, M, (int)Mp, 3 * D, (oess_stream_t)st);      // S = k q^T
    if (rc) return rc;
    OESS_KERNEL("nce_rows", st, k_nce_rows<<<(unsigned)M, 256, 0, st>>>(w.S, M, Mp, inv_T, w.lse, w.rowloss));
    OESS_KERNEL("nce_loss", st, k_nce_loss<<<1, 1024, 0, st>>>(w.rowloss, M, loss));
    const dim3 gg((unsigned)((Mp + 255) / 256), (unsigned)M), tg((unsigned)((Mp + 31) / 32), (unsigned)((D + 31) / 32));
    if (dk) {
        OESS_KERNEL("nce_grad_split", st, k_nce_grad_split<<<gg, 256, 0, st>>>(w.S, M, Mp, Mp, inv_T, inv_MT, w.lse, 0, w.G3));
        OESS_KERNEL("nce_split3_t", st, k_nce_split3_t<<<tg, 256, 0, st>>>(q, M, D, Mp, w.qT3));
        rc = oess_gemm_tf32(w.G3, w.qT3, nullptr, dk, M, D, (int)(3 * Mp), (oess_stream_t)st);                 // dk = G q
        if (rc) return rc;
    }
    if (dq) {
        OESS_KERNEL("nce_split3", st, k_nce_split3<<<sb, 256, 0, st>>>(q, M, D, Mp, 0, w.qa3));
        OESS_KERNEL("nce_split3", st, k_nce_split3<<<sb, 256, 0, st>>>(k, M, D, Mp, 1, w.kb3));
        rc = oess_gemm_tf32(w.qa3, w.kb3, nullptr, w.S, M, (int)Mp, 3 * D, (oess_stream_t)st);   // S^T = q k^T
        if (rc) return rc;
        OESS_KERNEL("nce_grad_split", st, k_nce_grad_split<<<gg, 256, 0, st>>>(w.S, M, Mp, Mp, inv_T, inv_MT, w.lse, 1, w.G3));
        OESS_KERNEL("nce_split3_t", st, k_nce_split3_t<<<tg, 256, 0, st>>>(k, M, D, Mp, w.kT3));
        rc = oess_gemm_tf32(w.G3, w.kT3, nullptr, dq, M, D, (int)(3 * Mp), (oess_stream_t)st);                 // dq = G^T k
        if (rc) return rc;
    }
    return OESS_OK;
}

}  // namespace oess

using namespace oess;

OESS_API int oess_infonce_ws_bytes(int64_t M, int D, size_t* ws_bytes) {
    if (!ws_bytes || M <= 0 || D <= 0) return OESS_E_ARG;
    *ws_bytes = nce_use_tc(M, D) ? nce_tc_carve(nullptr, M, D).bytes : nce_carve(nullptr, M, D).bytes;
    return OESS_OK;
}

OESS_API int oess_infonce(const float* k, const float* q, int64_t M, int D, float temperature, float* loss,
                          float* dk, float* dq, void* ws, size_t ws_bytes, oess_stream_t stream) {
    size_t need = 0;
    int rc = oess_infonce_ws_bytes(M, D, &need);
    if (rc) return rc;
    if (!k || !q || !loss) return OESS_E_ARG;
    if (!ws || ws_bytes < need) return OESS_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (nce_use_tc(M, D)) return run_infonce_tc(k, q, M, D, temperature, loss, dk, dq, nce_tc_carve(ws, M, D), st);
    const NceWs w = nce_carve(ws, M, D);
    // feature width D is zero-padded to 16 * DC columns (OpenESS uses D = 256)
    if (D <= 16) return run_infonce<1>(k, q, M, D, temperature, loss, dk, dq, w, st);
    if (D <= 32) return run_infonce<2>(k, q, M, D, temperature, loss, dk, dq, w, st);
    if (D <= 64) return run_infonce<4>(k, q, M, D, temperature, loss, dk, dq, w, st);
    if (D <= 128) return run_infonce<8>(k, q, M, D, temperature, loss, dk, dq, w, st);
    if (D <= 256) return run_infonce<16>(k, q, M, D, temperature, loss, dk, dq, w, st);
    return OESS_E_ARG;   // wider features are not used by any OpenESS configuration
}
