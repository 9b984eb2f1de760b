// pixel_linear.cu -- per-pixel linear maps on NCHW tensors with few channels (a 1x1 convolution whose weight fits
// in shared memory), the building block of the FUSED SemSegE2VID head:
//   models/style_networks.py:163-165   x32 -> conv1x1(32->256) -> conv1x1(256->512) -> conv(text_embeddings) = logits
// is a bias-affine chain with no non-linearity, so logits = W_eff x32 + b_eff with W_eff = T W512 W256 [K, 32].
// Computing it directly reads 128 B and writes 4K B per pixel instead of materialising the 256- and 512-channel
// full-resolution maps (2.3 GB + 4.6 GB at batch 8).  HBM-bound streaming kernels:
//   fwd        y[b, k, p] = bias[k] + sum_c W[k, c] x[b, c, p]
//   bwd input  same kernel with W^T and no bias
//   bwd weight dW[k, c] = sum_{b,p} dy[b, k, p] x[b, c, p],  db[k] = sum dy   (per-CTA partials, fixed-order reduce)
#include "common.cuh"

namespace oess {

constexpr int kPLMaxDim = 64;   // Cin, Cout <= 64 (weights <= 16 KB of shared memory)

template <int CIN_MAX>
__global__ void __launch_bounds__(256)
k_pixel_linear(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias, int B,
               int Cin, int Cout, int64_t HW, float* __restrict__ y) {
    __shared__ float s_w[kPLMaxDim * kPLMaxDim];
    __shared__ float s_b[kPLMaxDim];
    for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) s_w[i] = Wt[i];      // [Cout][Cin]
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int64_t total = (int64_t)B * HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t b = i / HW, p = i - b * HW;
        const float* xp = x + b * Cin * HW + p;
        float xv[CIN_MAX];
#pragma unroll
        for (int c = 0; c < CIN_MAX; ++c) xv[c] = (c < Cin) ? __ldcs(xp + (int64_t)c * HW) : 0.f;
        float* yp = y + b * Cout * HW + p;
        for (int k = 0; k < Cout; ++k) {
            float acc = s_b[k];
            const float* w = s_w + k * Cin;
#pragma unroll
            for (int c = 0; c < CIN_MAX; ++c)
                if (c < Cin) acc = fmaf(w[c], xv[c], acc);
            __stcs(yp + (int64_t)k * HW, acc);
        }
    }
}

// dW / db partials.  CTA loops over 128-pixel tiles staged in shared memory; thread (kg, cg) owns a 2 x 4 micro-tile.
constexpr int kPLTile = 64;

__global__ void __launch_bounds__(256)
k_pixel_linear_wgrad(const float* __restrict__ dy, const float* __restrict__ x, int B, int Cin, int Cout, int64_t HW,
                     float* __restrict__ part) {       // part[gridDim.x][Cout * Cin + Cout]
    __shared__ float s_d[kPLMaxDim][kPLTile + 1];
    __shared__ float s_x[kPLMaxDim][kPLTile + 1];
    const int tid = threadIdx.x;
    const int ncg = (Cin + 3) / 4, nkg = (Cout + 1) / 2;
    const int pairs = ncg * nkg;                       // <= 16 * 32 = 512 micro-tiles for 256 threads
    float acc[2][2][4];
    float accb[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            accb[u][a] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[u][a][c] = 0.f;
        }
    const int64_t total = (int64_t)B * HW;
    const int64_t ntiles = (total + kPLTile - 1) / kPLTile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t i0 = t * kPLTile;
        __syncthreads();
        for (int e = tid; e < (Cout + Cin) * kPLTile; e += 256) {
            const int row = e / kPLTile, pp = e - row * kPLTile;
            const int64_t i = i0 + pp;
            float v = 0.f;
            if (i < total) {
                const int64_t b = i / HW, p = i - b * HW;
                v = (row < Cout) ? __ldcs(dy + (b * Cout + row) * HW + p) : __ldcs(x + (b * Cin + (row - Cout)) * HW + p);
            }
            if (row < Cout) s_d[row][pp] = v; else s_x[row - Cout][pp] = v;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int pr = tid + u * 256;
            if (pr >= pairs) continue;
            const int kg = pr / ncg, cg = pr - kg * ncg;
            const int k0 = kg * 2, c0 = cg * 4;
            for (int pp = 0; pp < kPLTile; ++pp) {
                float dv[2], xv[4];
#pragma unroll
                for (int a = 0; a < 2; ++a) dv[a] = (k0 + a < Cout) ? s_d[k0 + a][pp] : 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) xv[c] = (c0 + c < Cin) ? s_x[c0 + c][pp] : 0.f;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    if (cg == 0) accb[u][a] += dv[a];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[u][a][c] = fmaf(dv[a], xv[c], acc[u][a][c]);
                }
            }
        }
    }
    float* po = part + (int64_t)blockIdx.x * (Cout * Cin + Cout);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int pr = tid + u * 256;
        if (pr >= pairs) continue;
        const int kg = pr / ncg, cg = pr - kg * ncg;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int k = kg * 2 + a;
            if (k >= Cout) continue;
            if (cg == 0) po[Cout * Cin + k] = accb[u][a];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (cg * 4 + c < Cin) po[k * Cin + cg * 4 + c] = acc[u][a][c];
        }
    }
}

__global__ void k_pixel_linear_wreduce(const float* __restrict__ part, int nparts, int n, float* __restrict__ dW,
                                       float* __restrict__ db, int nW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0;
    for (int s = 0; s < nparts; ++s) a += (double)part[(int64_t)s * n + i];
    if (i < nW) dW[i] = (float)a; else if (db) db[i - nW] = (float)a;
}

static inline int pl_wgrad_ctas() { return kNumSMs * 2; }

}  // namespace oess

using namespace oess;

OESS_API int oess_pixel_linear(const float* x, const float* W, const float* bias, int B, int Cin, int Cout, int64_t HW,
                               float* y, oess_stream_t stream) {
    if (B <= 0 || Cin <= 0 || Cout <= 0 || Cin > kPLMaxDim || Cout > kPLMaxDim || HW <= 0 || !x || !W || !y)
        return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)B * HW;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    if (Cin <= 8) { OESS_KERNEL("pixel_linear", st, k_pixel_linear<8><<<(unsigned)blocks, 256, 0, st>>>(x, W, bias, B, Cin, Cout, HW, y)); }
    else if (Cin <= 16) { OESS_KERNEL("pixel_linear", st, k_pixel_linear<16><<<(unsigned)blocks, 256, 0, st>>>(x, W, bias, B, Cin, Cout, HW, y)); }
    else if (Cin <= 32) { OESS_KERNEL("pixel_linear", st, k_pixel_linear<32><<<(unsigned)blocks, 256, 0, st>>>(x, W, bias, B, Cin, Cout, HW, y)); }
    else { OESS_KERNEL("pixel_linear", st, k_pixel_linear<64><<<(unsigned)blocks, 256, 0, st>>>(x, W, bias, B, Cin, Cout, HW, y)); }
    return OESS_OK;
}

OESS_API int oess_pixel_linear_wgrad_ws_bytes(int Cin, int Cout, size_t* ws_bytes) {
    if (!ws_bytes || Cin <= 0 || Cout <= 0 || Cin > kPLMaxDim || Cout > kPLMaxDim) return OESS_E_ARG;
    *ws_bytes = sizeof(float) * (size_t)pl_wgrad_ctas() * (Cout * Cin + Cout);
    return OESS_OK;
}

OESS_API int oess_pixel_linear_wgrad(const float* dy, const float* x, int B, int Cin, int Cout, int64_t HW, float* dW,
                                     float* db, void* ws, size_t ws_bytes, oess_stream_t stream) {
    size_t need = 0;
    int rc = oess_pixel_linear_wgrad_ws_bytes(Cin, Cout, &need);
    if (rc) return rc;
    if (B <= 0 || HW <= 0 || !dy || !x || !dW) return OESS_E_ARG;
    if (!ws || ws_bytes < need) return OESS_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = Cout * Cin + Cout;
    const int ctas = pl_wgrad_ctas();
    OESS_CUDA(cudaMemsetAsync(ws, 0, need, st));
    OESS_KERNEL("pixel_linear_wgrad", st, k_pixel_linear_wgrad<<<ctas, 256, 0, st>>>(dy, x, B, Cin, Cout, HW, (float*)ws));
    OESS_KERNEL("pixel_linear_wreduce", st, k_pixel_linear_wreduce<<<(n + 255) / 256, 256, 0, st>>>(
        (const float*)ws, ctas, n, dW, db, Cout * Cin));
    return OESS_OK;
}
