// tc_gemm.cu -- TF32 tensor-core GEMM for the dense contractions of the path that are plain matrix products:
//     C[M, N] = A[M, K] * B[N, K]^T (+ bias[N])          A, B, C row-major float32
// i.e. nn.Linear / a 1x1 convolution over channels-last pixels (models/image_model.py:121-124 decoder conv
// 2048 -> 256, models/maskclip_model.py ViT linears, models/style_networks.py:163-165 head convs).
//
// sm_100a structure: one CTA per 128 x BN output tile; warp 0 = TMA producer (cp.async.bulk.tensor into a
// 4-stage 128B-swizzled shared-memory ring), warp 1 = MMA issuer (one thread, tcgen05.mma.kind::tf32, accumulator
// in TMEM), warps 2..5 = epilogue (tcgen05.ld -> + bias -> global).  Operands are fp32 in memory; the tensor core
// reads them as TF32 (10-bit mantissa), accumulation is fp32 -- the same arithmetic class torch's default cuDNN /
// cuBLAS-TF32 convolution path uses on the reference's GPU run.  Stated tolerance: 2e-3 relative to |A||B|.
#include <stdlib.h>

#include "tc_common.cuh"

namespace oess {
namespace tc {

static EncodeTiledFn g_encode = nullptr;

EncodeTiledFn encode_tiled_fn() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
    }
    return g_encode;
}

static int make_tmap_f32_sw(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, CUtensorMapSwizzle sw, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32);

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    return make_tmap_f32_sw(m, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap_f32_sw(m, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap_f32_atom32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box) {
    return make_tmap_f32_sw(m, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

static int make_tmap_f32_sw(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, CUtensorMapSwizzle sw, CUtensorMapDataType dt) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

constexpr int kBM = 128;
constexpr int kGemmThreads = 192;

// nn.GELU() (erf form), the activation of mmcv's FFN in models/maskclip_model.py:507-513
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Ring depth: sized so that TWO CTAs share an SM (<= ~97 KB each): the TMEM epilogue of one tile (bias / GELU / residual,
// 128 x BN global stores) runs under the main loop of the other CTA's tile -- same idea as the ConvLSTM kernel.
template <int BN, int kStages>
struct GemmSmem {
    static constexpr int kABytes = kBM * kBlockK * 4;      // 16 KB
    static constexpr int kBBytes = BN * kBlockK * 4;
    static constexpr int kBytes = 1024 + kStages * (kABytes + kBBytes) + 256;
};

// MC: clusters of two CTAs with adjacent M tiles and the same N tile; each loads its own A tile and HALF of the B tile, which
// TMA multicasts into both CTAs' shared memory (a stage is released to BOTH producers: the commit arrives on the empty barrier
// of both CTAs, count 2).  Measured: no gain (122 vs 114 us at M = 8968, N = 2304, K = 768) -- the L2 -> SM read path is only
// 20 % utilised in this kernel (ncu), so halving the requests does not help; kept as OESS_GEMM=mc.
template <int BN, int kStages, bool MC>
__global__ void __launch_bounds__(kGemmThreads, (GemmSmem<BN, kStages>::kBytes <= 100 * 1024) ? 2 : 1)
k_gemm_tf32(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const float* __restrict__ bias, const float* residual, float* C, int64_t M, int N, int K, int act) {
    extern __shared__ uint8_t smem_raw[];
    using S = GemmSmem<BN, kStages>;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned tiles
    uint8_t* sA = base;
    uint8_t* sB = base + kStages * S::kABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kStages * S::kBBytes);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int kblocks = (K + kBlockK - 1) / kBlockK;
    const uint32_t crank = MC ? cluster_ctarank() : 0u;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MC ? 2 : 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();      // peers' barriers exist before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                  // ===== TMA producer =====
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], S::kABytes + S::kBBytes);
                tma_load_2d(sA + s * S::kABytes, &tmA, &full[s], kb * kBlockK, (int)m0);
                if (MC)
                    tma_load_2d_mc(sB + s * S::kBBytes + crank * (S::kBBytes / 2), &tmB, &full[s], kb * kBlockK,
                                   n0 + (int)crank * (BN / 2), (uint16_t)3);
                else
                    tma_load_2d(sB + s * S::kBBytes, &tmB, &full[s], kb * kBlockK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                  // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&full[s], (kb / kStages) & 1);
                tc_fence_after();
                const uint64_t da = umma_desc_k128(smem_u32(sA + s * S::kABytes));
                const uint64_t db = umma_desc_k128(smem_u32(sB + s * S::kBBytes));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)   // +32 bytes (>> 4 = 2) per K = 8 step inside the swizzle line
                    umma_tf32(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                if (MC) umma_commit_mc(&empty[s], (uint16_t)3);   // frees the stage in BOTH CTAs' producers' eyes
                else umma_commit(&empty[s]);              // frees the smem stage once these MMAs have read it
            }
            umma_commit(acc_full);                        // accumulator complete
        }
    } else {                                              // ===== epilogue: warps 2..5 =====
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int64_t row = m0 + q * 32 + lane;
        float* crow = C + row * (int64_t)N;
        const float* rrow = residual ? residual + row * (int64_t)N : nullptr;   // may alias C (same element, same thread)
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            const int col = n0 + c0;
            if (row < M && col < N) {
                if (col + 32 <= N && (N & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (bias) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + col + j);
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (act & 1) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
                        if (rrow) {
                            const float4 r = *reinterpret_cast<const float4*>(rrow + col + j);
                            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                        }
                        if (act & 2) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
                        *reinterpret_cast<float4*>(crow + col + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {        // fully unrolled + predicated: v[] stays in registers
                        if (col + j < N) {
                            float o = v[j] + (bias ? bias[col + j] : 0.0f);
                            if (act & 1) o = gelu_erf(o);
                            if (rrow) o += rrow[col + j];
                            crow[col + j] = (act & 2) ? rna_tf32(o) : o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();      // no CTA leaves while its peer can still write into it
    if (warp == 1) tmem_dealloc(tmem_acc, BN);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default when N % 4 == 0): one CTA per SM walks tiles blockIdx.x, + gridDim.x, ...; a 4 / 6 / 8-stage
// operand ring that keeps running across tile boundaries; TWO TMEM accumulators (2 x BN columns), so the MMA warp starts the
// next tile while the eight epilogue warps drain the previous one; each epilogue warp stages 32 x 32 blocks of its TMEM lane
// quarter in shared memory (the 128-byte swizzle pattern, conflict-free float4 stores) and writes them with its own TMA stores
// (cp.async.bulk.tensor ... .global.shared::cta), which also clip the ragged M / N edges.
template <int BN>
struct PGemmSmem {
    static constexpr int kStages = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
    static constexpr int kABytes = kBM * kBlockK * 4;      // 16 KB
    static constexpr int kBBytes = BN * kBlockK * 4;
    static constexpr int kCBytes = kBM * 32 * 4;           // one staged 128 x 32 output block
    static constexpr int kBytes = 1024 + kStages * (kABytes + kBBytes) + 2 * kCBytes + 256;
};

constexpr int kPGemmThreads = 320;          // TMA producer warp, MMA warp, eight epilogue warps

template <int BN>
__global__ void __launch_bounds__(kPGemmThreads, 1)
k_gemm_tf32_persist(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const float* residual,
                    int64_t M, int N, int K, int act, int tiles_m, int tiles) {
    extern __shared__ uint8_t smem_raw[];
    using S = PGemmSmem<BN>;
    constexpr int kStages = S::kStages;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = sA + kStages * S::kABytes;
    uint8_t* sC = sB + kStages * S::kBBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + 2 * S::kCBytes);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;                  // [2]
    uint64_t* acc_empty = acc_full + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);                   // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {                                      // ===== TMA producer (whole warp converged, one lane issues) =====
        uint32_t s = 0, ph = 1;                           // ring slot and the parity to wait for on its empty barrier
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m0 = (tile % tiles_m) * kBM, n0 = (tile / tiles_m) * BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], S::kABytes + S::kBBytes);
                    tma_load_2d(sA + s * S::kABytes, &tmA, &full[s], kb * kBlockK, m0);
                    tma_load_2d(sB + s * S::kBBytes, &tmB, &full[s], kb * kBlockK, n0);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {                               // ===== MMA issuer (whole warp converged, one lane issues) =====
        constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
        uint32_t s = 0, ph = 0, lt = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
            const uint32_t buf = lt & 1;
            mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem_acc + buf * BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_k128(smem_u32(sA + s * S::kABytes));
                    const uint64_t db = umma_desc_k128(smem_u32(sB + s * S::kBBytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_tf32(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {                                              // ===== epilogue: warps 2..9 =====
        // Two warps per TMEM lane quarter (one per half of the tile's columns): every SM sub-partition has two epilogue warps to
        // switch between.  A warp owns its 32 rows x BN / 2 columns end to end -- its own 32 x 32 staging block (4 KB) and its
        // own TMA stores, no barrier between the warps -- and keeps the next 32 columns' tcgen05.ld in flight while it
        // finishes the current ones.
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        constexpr int NCH = BN / 64;                      // 32-column chunks per warp and tile
        uint8_t* stage = sC + (warp - 2) * 4096;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
            const int m0 = (tile % tiles_m) * kBM, n0 = (tile / tiles_m) * BN + half * (BN / 2);
            const uint32_t buf = lt & 1;
            mbar_wait(&acc_full[buf], (lt >> 1) & 1);
            tc_fence_after();
            const int64_t row = (int64_t)m0 + q * 32 + lane;
            const float* rrow = (residual && row < M) ? residual + row * (int64_t)N : nullptr;   // may alias C
            const uint32_t t0 = tmem_acc + buf * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2));
            float v[2][32];
            tmem_ld32_nowait(t0, v[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                tmem_ld_wait();
                if (ch + 1 < NCH) tmem_ld32_nowait(t0 + (uint32_t)(ch + 1) * 32, v[(ch + 1) & 1]);
                if (ch + 1 == NCH) {                      // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                const int col = n0 + ch * 32;
                if (col >= N || m0 + q * 32 >= M) continue;   // warp-uniform (ragged last tiles)
                const bool full32 = col + 32 <= N;
                float (&w)[32] = v[ch & 1];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (full32 || col + j + 4 <= N) {     // N % 4 == 0: whole quads are inside or outside
                        if (bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + j));
                            w[j] += b.x; w[j + 1] += b.y; w[j + 2] += b.z; w[j + 3] += b.w;
                        }
                        if (act & 1) {
                            w[j] = gelu_erf(w[j]); w[j + 1] = gelu_erf(w[j + 1]); w[j + 2] = gelu_erf(w[j + 2]); w[j + 3] = gelu_erf(w[j + 3]);
                        }
                        if (rrow) {
                            const float4 rr = *reinterpret_cast<const float4*>(rrow + col + j);
                            w[j] += rr.x; w[j + 1] += rr.y; w[j + 2] += rr.z; w[j + 3] += rr.w;
                        }
                        if (act & 2) {
                            w[j] = rna_tf32(w[j]); w[j + 1] = rna_tf32(w[j + 1]); w[j + 2] = rna_tf32(w[j + 2]); w[j + 3] = rna_tf32(w[j + 3]);
                        }
                    }
                }
                if (elect_one()) tma_store_wait_read<0>();   // this warp's previous store has read the staging block
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4)           // 128-byte swizzle: 16-byte chunk j / 4 of row `lane` -> chunk (j / 4) ^ (lane & 7)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) =
                        make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {                        // always the same lane: bulk async-groups are per thread
                    tma_store_2d(&tmC, stage, col, m0 + q * 32);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, 2 * BN);
}

template <int BN>
static int launch_gemm_persist(const float* A, const float* B, const float* bias, const float* residual, float* C, int64_t M,
                               int N, int K, int act, cudaStream_t st) {
    CUtensorMap tmA, tmB, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)K * 4};
    const uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)K * 4};
    const uint64_t dC[2] = {(uint64_t)N, (uint64_t)M}, sC[1] = {(uint64_t)N * 4};
    const uint32_t bA[2] = {kBlockK, kBM}, bB[2] = {kBlockK, (uint32_t)BN}, bC[2] = {32, 32};
    int rc = make_tmap_f32(&tmA, A, 2, dA, sA, bA);
    if (rc) return rc;
    rc = make_tmap_f32(&tmB, B, 2, dB, sB, bB);
    if (rc) return rc;
    rc = make_tmap_f32(&tmC, C, 2, dC, sC, bC);
    if (rc) return rc;
    auto kern = k_gemm_tf32_persist<BN>;
    OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PGemmSmem<BN>::kBytes));
    const int tiles_m = (int)((M + kBM - 1) / kBM);
    const int64_t tiles = (int64_t)tiles_m * ((N + BN - 1) / BN);
    if (tiles >= (1ll << 31)) return OESS_E_RANGE;
    const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
    OESS_KERNEL("tc_gemm_tf32", st, kern<<<grid, kPGemmThreads, PGemmSmem<BN>::kBytes, st>>>(
        tmA, tmB, tmC, bias, residual, M, N, K, act, tiles_m, (int)tiles));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel for 256-wide tiles (cta_group::2; default for N > 128 when there are at least as many
// 256 x 256 tiles as CTA pairs, OESS_GEMM_2SM=0: off).  The two CTAs of a cluster compute ONE 256 x 256 tile per step: each stages
// its own 128 rows of A (16 KB) and HALF of the B tile (16 KB) per K block, so the ring holds SIX K blocks instead of four and
// every SM reads 8 KB instead of 12 KB of shared memory per K = 8 MMA (the TF32 product at BN = 256 otherwise reads + fills
// 192 B / clk of shared memory per SM); the leader issues `tcgen05.mma.cta_group::2` with M = 256 and commits to both CTAs'
// barriers; every CTA's epilogue drains its own 128 TMEM lanes exactly as in k_gemm_tf32_persist.
// BN = 192 as well as 256: the tile width is chosen per product so that the last wave of pair tiles is as full as possible
// (N = 768 on M = 8 968: 108 tiles of 256 x 256 are 1.46 waves of 74 pairs, 144 tiles of 256 x 192 are 1.95).
constexpr int kG2Stages = 6;
constexpr int kG2ABytes = kBM * kBlockK * 4;                // 16 KB
constexpr int kG2CBytes = kBM * 32 * 4;
template <int BN>
struct G2Smem {
    static constexpr int kBBytes = (BN / 2) * kBlockK * 4;  // 16 / 12 KB: this CTA's half of the B tile
    static constexpr int kBytes = 1024 + kG2Stages * (kG2ABytes + kBBytes) + 2 * kG2CBytes + 256;
};

template <int BN>
__global__ void __launch_bounds__(kPGemmThreads, 1)
k_gemm_tf32_p2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const float* residual, int64_t M, int N,
               int K, int act, int tiles_m, int tiles) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kG2BBytes = G2Smem<BN>::kBBytes;
    constexpr uint32_t kAccStride = 256;                   // TMEM columns between the two accumulators (512 allocated)
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = sA + kG2Stages * kG2ABytes;
    uint8_t* sC = sB + kG2Stages * kG2BBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + 2 * kG2CBytes);
    uint64_t* empty = full + kG2Stages;
    uint64_t* acc_full = empty + kG2Stages;                // [2]
    uint64_t* acc_empty = acc_full + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (K + kBlockK - 1) / kBlockK;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < kG2Stages; ++s) {
            mbar_init(&full[s], 1);                        // the leader's own arrive.expect_tx (both CTAs' bytes)
            mbar_init(&empty[s], 1);                       // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);                    // the leader's multicast commit
            mbar_init(&acc_empty[b], 16);                  // eight epilogue warps in each CTA (used in the leader only)
        }
        mbar_fence_init();
    }
    cluster_sync_all();                                    // barriers of both CTAs exist before TMEM allocation / any remote arrive
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const int tile0 = (int)(blockIdx.x >> 1), tstep = (int)(gridDim.x >> 1);

    if (warp == 0) {                                       // ===== TMA producer (both CTAs) =====
        uint32_t s = 0, ph = 1;
        for (int tile = tile0; tile < tiles; tile += tstep) {
            const int m0 = (tile % tiles_m) * 256 + (int)crank * kBM, n0 = (tile / tiles_m) * BN + (int)crank * (BN / 2);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[s], 2 * (kG2ABytes + kG2BBytes));
                    tma_load_2d_2sm(sA + s * kG2ABytes, &tmA, &full[s], kb * kBlockK, m0);
                    tma_load_2d_2sm(sB + s * kG2BBytes, &tmB, &full[s], kb * kBlockK, n0);
                }
                __syncwarp();
                if (++s == kG2Stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {                                // ===== MMA issuer (leader CTA only) =====
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_tf32(256, BN);
            uint32_t s = 0, ph = 0, lt = 0;
            for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
                const uint32_t buf = lt & 1;
                mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_acc + buf * kAccStride;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = umma_desc_k128(smem_u32(sA + s * kG2ABytes));
                        const uint64_t db = umma_desc_k128(smem_u32(sB + s * kG2BBytes));
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)
                            umma_tf32_2sm(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty[s], (uint16_t)3);
                    }
                    __syncwarp();
                    if (++s == kG2Stages) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(&acc_full[buf], (uint16_t)3);
                __syncwarp();
            }
        }
    } else {                                               // ===== epilogue: warps 2..9 of both CTAs, own 128 rows each =====
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        constexpr int NCH = BN / 64;
        uint8_t* stage = sC + (warp - 2) * 4096;
        uint32_t lt = 0;
        for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
            const int m0 = (tile % tiles_m) * 256 + (int)crank * kBM, n0 = (tile / tiles_m) * BN + half * (BN / 2);
            const uint32_t buf = lt & 1;
            mbar_wait(&acc_full[buf], (lt >> 1) & 1);
            tc_fence_after();
            const int64_t row = (int64_t)m0 + q * 32 + lane;
            const float* rrow = (residual && row < M) ? residual + row * (int64_t)N : nullptr;   // may alias C
            const uint32_t t0 = tmem_acc + buf * kAccStride + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2));
            float v[2][32];
            tmem_ld32_nowait(t0, v[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                tmem_ld_wait();
                if (ch + 1 < NCH) tmem_ld32_nowait(t0 + (uint32_t)(ch + 1) * 32, v[(ch + 1) & 1]);
                if (ch + 1 == NCH) {                       // accumulator fully read: tell the LEADER's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
                }
                const int col = n0 + ch * 32;
                if (col >= N || m0 + q * 32 >= M) continue;   // warp-uniform (ragged last tiles)
                const bool full32 = col + 32 <= N;
                float (&w)[32] = v[ch & 1];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (full32 || col + j + 4 <= N) {
                        if (bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + j));
                            w[j] += b.x; w[j + 1] += b.y; w[j + 2] += b.z; w[j + 3] += b.w;
                        }
                        if (act & 1) {
                            w[j] = gelu_erf(w[j]); w[j + 1] = gelu_erf(w[j + 1]); w[j + 2] = gelu_erf(w[j + 2]); w[j + 3] = gelu_erf(w[j + 3]);
                        }
                        if (rrow) {
                            const float4 rr = *reinterpret_cast<const float4*>(rrow + col + j);
                            w[j] += rr.x; w[j + 1] += rr.y; w[j + 2] += rr.z; w[j + 3] += rr.w;
                        }
                        if (act & 2) {
                            w[j] = rna_tf32(w[j]); w[j + 1] = rna_tf32(w[j + 1]); w[j + 2] = rna_tf32(w[j + 2]); w[j + 3] = rna_tf32(w[j + 3]);
                        }
                    }
                }
                if (elect_one()) tma_store_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) =
                        make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {
                    tma_store_2d(&tmC, stage, col, m0 + q * 32);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();                                    // no CTA leaves (or frees TMEM) while its peer still works
    if (warp == 1) tmem_dealloc_2sm(tmem_acc, 512);
}

template <int BN>
static int launch_gemm_p2(const float* A, const float* B, const float* bias, const float* residual, float* C, int64_t M, int N,
                          int K, int act, cudaStream_t st) {
    CUtensorMap tmA, tmB, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)K * 4};
    const uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)K * 4};
    const uint64_t dC[2] = {(uint64_t)N, (uint64_t)M}, sC[1] = {(uint64_t)N * 4};
    const uint32_t bA[2] = {kBlockK, kBM}, bB[2] = {kBlockK, (uint32_t)(BN / 2)}, bC[2] = {32, 32};
    int rc = make_tmap_f32(&tmA, A, 2, dA, sA, bA);
    if (rc) return rc;
    rc = make_tmap_f32(&tmB, B, 2, dB, sB, bB);
    if (rc) return rc;
    rc = make_tmap_f32(&tmC, C, 2, dC, sC, bC);
    if (rc) return rc;
    auto kern = k_gemm_tf32_p2<BN>;
    constexpr int kG2Smem = G2Smem<BN>::kBytes;
    OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2Smem));
    const int tiles_m = (int)((M + 255) / 256);
    const int64_t tiles = (int64_t)tiles_m * ((N + BN - 1) / BN);
    if (tiles >= (1ll << 30)) return OESS_E_RANGE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * tiles < kNumSMs ? 2 * tiles : (kNumSMs & ~1)));
    cfg.blockDim = dim3(kPGemmThreads);
    cfg.dynamicSmemBytes = kG2Smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OESS_KERNEL("tc_gemm_tf32", st, cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, bias, residual, M, N, K, act, tiles_m, (int)tiles));
    return 0;
}

template <int BN, int kStages, bool MC>
static int launch_gemm(const float* A, const float* B, const float* bias, const float* residual, float* C, int64_t M, int N,
                       int K, int act, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)K * 4};
    const uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)K * 4};
    const uint32_t bA[2] = {kBlockK, kBM}, bB[2] = {kBlockK, (uint32_t)(MC ? BN / 2 : BN)};
    int rc = make_tmap_f32(&tmA, A, 2, dA, sA, bA);
    if (rc) return rc;
    rc = make_tmap_f32(&tmB, B, 2, dB, sB, bB);
    if (rc) return rc;
    auto kern = k_gemm_tf32<BN, kStages, MC>;
    OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN, kStages>::kBytes));
    unsigned gx = (unsigned)((M + kBM - 1) / kBM);
    if (MC) gx = (gx + 1) & ~1u;                           // whole clusters: a padding CTA loads zeros and stores nothing
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, (unsigned)((N + BN - 1) / BN));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = GemmSmem<BN, kStages>::kBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MC ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OESS_KERNEL("tc_gemm_tf32", st, cudaLaunchKernelEx(&cfg, kern, tmA, tmB, bias, residual, C, M, N, K, act));
    return 0;
}

}  // namespace tc
}  // namespace oess

using namespace oess;

OESS_API int oess_gemm_tf32(const float* A, const float* B, const float* bias, float* C, int64_t M, int N, int K,
                            oess_stream_t stream) {
    return oess_gemm_tf32_ex(A, B, bias, nullptr, C, M, N, K, 0, stream);
}

OESS_API int oess_gemm_tf32_ex(const float* A, const float* B, const float* bias, const float* residual, float* C, int64_t M,
                               int N, int K, int act, oess_stream_t stream) {
    if (M < 0 || N <= 0 || K <= 0 || act < 0 || act > 3) return OESS_E_ARG;
    if (M == 0) return OESS_OK;
    if (!A || !B || !C) return OESS_E_ARG;
    // TMA: 16-byte aligned bases and row strides
    if ((K & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || ((uintptr_t)bias & 15) ||
        ((uintptr_t)residual & 15)) return OESS_E_ARG;
    if (M >= (1ll << 31)) return OESS_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    // default: two CTAs per SM (ring depth 2-4).  OESS_GEMM=deep: one CTA per SM with a 4-stage ring; OESS_GEMM=mc: two CTAs
    // per SM in clusters of two with the B tile multicast (measured: no gain, see the kernel comment)
    static const int variant = [] {
        const char* e = getenv("OESS_GEMM");
        return !e ? 3 : (e[0] == 'd' ? 0 : (e[0] == 'm' ? 2 : (e[0] == 't' ? 1 : 3)));   // t: one tile per CTA, two CTAs per SM
    }();
    if (variant == 3 && (N & 3) == 0) {
        // persistent kernel; tile width: 256 unless that leaves most SMs without a tile (tall-skinny products)
        const int64_t mt = (M + 127) / 128;
        int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
        while (bn > 64 && mt * ((N + bn - 1) / bn) * 5 < (int64_t)kNumSMs * 3) bn >>= 1;
        static const bool sm2 = [] { const char* e = getenv("OESS_GEMM_2SM"); return !e || e[0] != '0'; }();
        if (bn == 256 && sm2 && K > 2 * tc::kBlockK && ((M + 255) / 256) * ((N + 255) / 256) >= kNumSMs / 2) {   // K <= 64: store-bound, the pair only adds cluster latency (measured)
            // tile width 256 or 192: whichever needs fewer (waves of 74 pair tiles) x (columns per tile)
            static const bool w192 = [] { const char* e = getenv("OESS_GEMM_192"); return !e || e[0] != '0'; }();
            const int64_t pairs = kNumSMs / 2, tm2 = (M + 255) / 256;
            const int64_t c256 = ((tm2 * ((N + 255) / 256) + pairs - 1) / pairs) * 256;
            const int64_t c192 = ((tm2 * ((N + 191) / 192) + pairs - 1) / pairs) * 192;
            // (a 192-wide tile costs more per column -- the A fill per K block is the same -- so it has to save an eighth: measured)
            if (w192 && c192 * 8 < c256 * 7) return tc::launch_gemm_p2<192>(A, B, bias, residual, C, M, N, K, act, st);
            return tc::launch_gemm_p2<256>(A, B, bias, residual, C, M, N, K, act, st);
        }
        if (bn == 256) return tc::launch_gemm_persist<256>(A, B, bias, residual, C, M, N, K, act, st);
        if (bn == 128) return tc::launch_gemm_persist<128>(A, B, bias, residual, C, M, N, K, act, st);
        return tc::launch_gemm_persist<64>(A, B, bias, residual, C, M, N, K, act, st);
    }
    // fewer tiles than SMs: a second resident CTA has nothing to overlap with, the deeper ring hides the load latency instead
    const int64_t tiles = ((M + 127) / 128) * (int64_t)((N + (N > 128 ? 255 : (N > 64 ? 127 : 63))) / (N > 128 ? 256 : (N > 64 ? 128 : 64)));
    if (variant == 0 || (variant != 2 && tiles <= kNumSMs)) {
        // tall-skinny products (e.g. the InfoNCE gradient G q: M = 3 200, N = 256, K = 9 600 -> 25 tiles of 128 x 256): narrower
        // N tiles put more SMs to work; the extra A-tile reads hit L2
        const int64_t mt = (M + 127) / 128;
        const bool few = variant != 0 && tiles * 5 < kNumSMs * 3;
        const int bn = !few ? (N > 128 ? 256 : (N > 64 ? 128 : 64))
                            : ((N > 128 && mt * ((N + 127) / 128) * 5 >= kNumSMs * 3) ? 128 : 64);
        if (bn == 256) return tc::launch_gemm<256, 4, false>(A, B, bias, residual, C, M, N, K, act, st);
        if (bn == 128) return tc::launch_gemm<128, 4, false>(A, B, bias, residual, C, M, N, K, act, st);
        return tc::launch_gemm<64, 4, false>(A, B, bias, residual, C, M, N, K, act, st);
    }
    if (variant != 2) {
        if (N > 128) return tc::launch_gemm<256, 2, false>(A, B, bias, residual, C, M, N, K, act, st);
        if (N > 64) return tc::launch_gemm<128, 3, false>(A, B, bias, residual, C, M, N, K, act, st);
        return tc::launch_gemm<64, 4, false>(A, B, bias, residual, C, M, N, K, act, st);
    }
    if (N > 128) return tc::launch_gemm<256, 2, true>(A, B, bias, residual, C, M, N, K, act, st);
    if (N > 64) return tc::launch_gemm<128, 3, true>(A, B, bias, residual, C, M, N, K, act, st);
    return tc::launch_gemm<64, 4, true>(A, B, bias, residual, C, M, N, K, act, st);
}
