// tc_wgrad.cu -- weight gradient of a stride-1 convolution as a tcgen05 split-K GEMM over pixels (TF32 operands, fp32
// accumulate), channels-last tensors:
//     dW[co, ci, ky, kx] = sum_{b, y, x} dY[b, y, x, co] * X[b, y + ky d - p, x + kx d - p, ci]
// Completes the convolution triple on the tensor cores (forward = tc_conv.cu, backward-data = the forward kernel on
// rotated weights, ops.conv2d_dgrad_pack) for the trainable convolutions of the path.
//
// GEMM view per filter row ky: M = 128 output channels, N = NT input channels, K = pixels.  With channels-last tensors
// the channel index is the contiguous one, so BOTH operands are MN-major: one 128-byte shared-memory line holds 32
// consecutive channels of ONE pixel.  A K block = 32 consecutive x of one image row:
//   A tile = 4 TMA boxes {32 co, 32 x, 1 y, 1 b} of dY (one per 32-channel group, 4 KB each, LBO = 4 KB apart),
//   B tile (one per kx) = NT / 32 boxes {32 ci, 32 x, 1 y, 1 b} of X at x + kx d - p, y + ky d - p.
// The shifts are on NON-innermost TMA coordinates (no 16-byte alignment constraint; a plane-major formulation would need
// x +- 1 element on the innermost coordinate, which TMA rejects), out-of-image coordinates are zero-filled = zero padding.
// The KW taps of a filter row share the A tile and accumulate into KW TMEM accumulators.
// Split-K: every CTA takes a contiguous range of the B * Ho image rows and adds its partial [128, NT, KW] result into dW
// with red.global.add.f32 (dW is zeroed inside the entry point).
#include "tc_common.cuh"

namespace oess {
namespace tc {

constexpr int kWStages = 3;
constexpr int kWGroupBytes = 32 * kBlockK * 4;            // one 32-channel group x 32 pixels = 4 KB
constexpr int kWABytes = 4 * kWGroupBytes;                // M = 128

struct WgradArgs {
    int Ho, Wo, Cin, Cout, KH, KW, pad, dil, NT, rows_total, rows_per_cta, xchunks;
};

__global__ void __launch_bounds__(192, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, float* __restrict__ dW,
           const WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int ngroups = a.NT / 32;
    const int bBytes = ngroups * kWGroupBytes;
    const int stageBytes = kWABytes + a.KW * bBytes;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + kWStages * stageBytes);
    uint64_t* empty = full + kWStages;
    uint64_t* acc_full = empty + kWStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tiles = (a.Cin + a.NT - 1) / a.NT;
    const int ci0 = (blockIdx.y % ci_tiles) * a.NT, co0 = (blockIdx.y / ci_tiles) * 128;
    const int ky = blockIdx.z;
    const int r_lo = blockIdx.x * a.rows_per_cta, r_hi = min(r_lo + a.rows_per_cta, a.rows_total);
    const int kblocks = max(r_hi - r_lo, 0) * a.xchunks;
    const uint32_t tmem_cols = 512;                       // KW * NT <= 512, power of two for the allocator

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (kblocks > 0) {
        if (warp == 0) {                                  // ===== TMA producer (whole warp converged, one lane issues) =====
            int r = r_lo, xc = 0;
            int b = r / a.Ho, y = r - b * a.Ho;
            uint32_t s = 0, ph = 1;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], (uint32_t)stageBytes);
                    uint8_t* st = base + s * stageBytes;
                    for (int gco = 0; gco < 4; ++gco)
                        tma_load_4d(st + gco * kWGroupBytes, &tmDY, &full[s], co0 + gco * 32, xc * kBlockK, y, b);
                    for (int kx = 0; kx < a.KW; ++kx)
                        for (int gci = 0; gci < ngroups; ++gci)
                            tma_load_4d(st + kWABytes + kx * bBytes + gci * kWGroupBytes, &tmX, &full[s], ci0 + gci * 32,
                                        xc * kBlockK + kx * a.dil - a.pad, y + ky * a.dil - a.pad, b);
                }
                __syncwarp();
                if (++xc == a.xchunks) { xc = 0; if (++y == a.Ho) { y = 0; ++b; } }
                if (++s == (uint32_t)kWStages) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1) {                           // ===== MMA issuer (whole warp converged, one lane issues) =====
            const uint32_t idesc = umma_idesc_tf32_mn(128, a.NT);
            uint32_t s = 0, ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    uint8_t* st = base + s * stageBytes;
                    const uint64_t da = umma_desc_mn128(smem_u32(st), kWGroupBytes);
                    for (int kx = 0; kx < a.KW; ++kx) {
                        const uint64_t db = umma_desc_mn128(smem_u32(st + kWABytes + kx * bBytes), kWGroupBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)    // 8 pixels = 8 lines = 1024 bytes (>> 4 = 64) per step
                            umma_tf32(tmem_acc + (uint32_t)(kx * a.NT), da + 64 * k, db + 64 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == (uint32_t)kWStages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
        } else {                                          // ===== epilogue: warps 2..5 =====
            const int q = warp & 3;
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int co = co0 + q * 32 + lane;
            const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
            for (int kx = 0; kx < a.KW; ++kx) {
#pragma unroll 1
                for (int c0 = 0; c0 < a.NT; c0 += 16) {
                    float v[16];
                    tmem_ld16_nowait(trow + (uint32_t)(kx * a.NT + c0), v);
                    tmem_ld_wait();
                    if (co < a.Cout) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int ci = ci0 + c0 + j;
                            if (ci < a.Cin)
                                atomicAdd(dW + (((int64_t)co * a.Cin + ci) * a.KH + ky) * a.KW + kx, v[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, tmem_cols);
}

}  // namespace tc
}  // namespace oess

using namespace oess;

// x: [B, H, W, Cin] and dy: [B, Ho, Wo, Cout] CHANNELS-LAST float32, stride-1 convolution with `pad` / `dil`
// (Ho = H + 2 pad - dil (KH - 1), same for W); dW: [Cout, Cin, KH, KW] (torch layout), overwritten.
// Cin % 4 == 0 and Cout % 4 == 0 (16-byte pixel strides for TMA), KW <= 5.
OESS_API int oess_conv2d_wgrad_nhwc_tf32(const float* x, const float* dy, float* dW, int B, int H, int W, int Cin, int Cout,
                                         int KH, int KW, int pad, int dil, oess_stream_t stream) {
    if (!x || !dy || !dW || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || KW > 5 || KH > 65535 ||
        pad < 0 || dil <= 0)
        return OESS_E_ARG;
    const int Ho = H + 2 * pad - dil * (KH - 1), Wo = W + 2 * pad - dil * (KW - 1);
    if (Ho <= 0 || Wo <= 0 || (Cin & 3) || (Cout & 3)) return OESS_E_ARG;
    if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dW) & 15) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // N tile: KW accumulators of NT columns must fit 512 TMEM columns; NT a multiple of 32 (whole channel groups)
    int NT = (512 / KW) >= 256 ? 256 : ((512 / KW) >= 128 ? 128 : 64);
    const int cin32 = ((Cin + 31) / 32) * 32;
    if (NT > cin32) NT = cin32;
    while (1024 + tc::kWStages * (tc::kWABytes + KW * (NT / 32) * tc::kWGroupBytes) + 256 > 227 * 1024 && NT > 32) NT -= 32;
    const int smem = 1024 + tc::kWStages * (tc::kWABytes + KW * (NT / 32) * tc::kWGroupBytes) + 256;
    if (smem > 227 * 1024) return OESS_E_ARG;
    CUtensorMap tmDY, tmX;
    {
        const uint64_t d[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
        const uint64_t s[3] = {(uint64_t)Cout * 4, (uint64_t)Wo * Cout * 4, (uint64_t)Ho * Wo * Cout * 4};
        const uint32_t bx[4] = {32, tc::kBlockK, 1, 1};
        int rc = tc::make_tmap_f32_atom32(&tmDY, dy, 4, d, s, bx);
        if (rc) return rc;
    }
    {
        const uint64_t d[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t s[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
        const uint32_t bx[4] = {32, tc::kBlockK, 1, 1};
        int rc = tc::make_tmap_f32_atom32(&tmX, x, 4, d, s, bx);
        if (rc) return rc;
    }
    OESS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)Cout * Cin * KH * KW, st));
    const int co_tiles = (Cout + 127) / 128, ci_tiles = (Cin + NT - 1) / NT;
    const int rows_total = B * Ho;
    // split K so that ~2 waves of CTAs exist, but every CTA keeps >= 8 image rows (amortises its 128 x NT x KW atomics)
    int splits = (2 * kNumSMs + co_tiles * ci_tiles * KH - 1) / (co_tiles * ci_tiles * KH);
    if (splits < 1) splits = 1;
    int rpc = (rows_total + splits - 1) / splits;
    if (rpc < 8) rpc = rows_total < 8 ? rows_total : 8;
    splits = (rows_total + rpc - 1) / rpc;
    tc::WgradArgs a{Ho, Wo, Cin, Cout, KH, KW, pad, dil, NT, rows_total, rpc, (Wo + tc::kBlockK - 1) / tc::kBlockK};
    OESS_CUDA(cudaFuncSetAttribute(tc::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const dim3 grid((unsigned)splits, (unsigned)(co_tiles * ci_tiles), (unsigned)KH);
    OESS_KERNEL("tc_conv2d_wgrad", st, tc::k_wgrad_tc<<<grid, 192, smem, st>>>(tmDY, tmX, dW, a));
    return OESS_OK;
}
