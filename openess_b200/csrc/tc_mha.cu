// tc_mha.cu -- multi-head attention forward of the MaskCLIP ViT-B/16 (models/maskclip_model.py:519-541, mmcv
// MultiheadAttention = nn.MultiheadAttention) on the tensor cores: softmax(Q K^T / sqrt(64)) V per (sample, head), flash
// style, both products as tcgen05.mma.kind::tf32 with the accumulators in TMEM.
//
// One CTA = 128 queries of one (sample, head); 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warps 2..5 = softmax (thread = query row = TMEM lane).  Per 64-key tile:
//   TMA   K tile [64 keys x 64 d] (two K-major 128B-swizzled blocks), V tile as FOUR {32 d, 32 keys} boxes in the MN-major
//         SWIZZLE_128B_ATOM_32B layout (V is [keys, d] in memory: the contraction index is the strided one);
//   MMA   S[128 x 64] = Q K^T            (8 x tcgen05.mma M = 128, N = 64, K = 8; Q stays in shared memory for all tiles)
//   warps S -> registers (tcgen05.ld), scale, mask the tail keys, online max / sum, P = exp2(s - m) rounded to TF32 and
//         written to shared memory in the K-major swizzled layout the tensor core reads as the A operand
//   MMA   O_tile[128 x 64] = P V         (A K-major from the softmax threads, B MN-major from TMA)
//   warps O = O * exp2(m_old - m_new) + O_tile in registers (one output row per thread)
// K and V have separate full / empty barriers, so K of tile j + 1 streams in under the softmax of tile j and V under the
// next S product; S is double-buffered in TMEM (S of tile j + 1 is issued before P V of tile j); two CTAs share an SM
// (97 KB shared memory, 256 TMEM columns each) and fill each other's bubbles.
// Measured (B = 8, T = 1121, 12 heads, per layer): 4 softmax warps + single S buffer 131 us; 8 softmax warps (two threads per
// row) 132 us; + double-buffered S 132-135 us -- neither the per-thread softmax chain nor the S hand-off is what bounds it.
// Tail handling: the tensor maps are 3-D {3 D, T, B}; rows past T are zero-filled by the TMA unit.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace oess {
namespace tc {

constexpr int kMhaQ = 128, kMhaK = 64, kMhaD = 64;
constexpr int kQBytes = kMhaQ * kMhaD * 4;                 // 32 KB: 2 K blocks of [128 x 32]
constexpr int kKBytes = kMhaK * kMhaD * 4;                 // 16 KB: 2 K blocks of [64 x 32]
constexpr int kVBytes = kMhaK * kMhaD * 4;                 // 16 KB: [2 key blocks][2 d groups][32 lines x 128 B]
constexpr int kPBytes = kMhaQ * kMhaK * 4;                 // 32 KB: 2 K blocks of [128 x 32]
constexpr int kMhaSmemBytes = 1024 + kQBytes + kKBytes + kVBytes + kPBytes + 128 + 1024;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// kind::tf32, A K-major, B MN-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32_bmn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// p >= 0: adding half a TF32 ulp to the bit pattern makes the tensor core's truncation a round-to-nearest (one IADD; the
// sum of the row uses the same value the MMA sees up to the dropped bits)
__device__ __forceinline__ float round_up_tf32(float p) { return __uint_as_float((__float_as_uint(p) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// SW = softmax warps (4 or 8).  With 8, two threads share a query row (32 of the tile's 64 keys / 32 of the 64 output
// columns each; row maxima and sums meet in shared memory): half the dependent instruction chain per tile and twice the
// warps per SM sub-partition to hide it -- the kernel is bound by that chain, not by the tensor pipe.
// FAST_EX2: ex2.approx.ftz.f32 instead of exp2f() (whose denormal-range handling costs ~5 extra instructions per element:
// the ncu instruction mix of this kernel is 15 % FMUL / 7 % FSETP / 4 % FSEL around 4 % MUFU.EX2, and issue slots -- half
// of them also burnt by barrier polling -- are what it runs out of).  Arguments are <= 0 here, results in [0, 1]; values
// below 2^-126 flush to zero.  Default since round 2 (parity suite green, 1.544 vs 1.58 ms per ViT forward); OESS_MHA_EX2=exact
// selects exp2f().
template <bool FAST>
__device__ __forceinline__ float ex2(float x) {
    if (FAST) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    return exp2f(x);
}

template <int SW, bool FAST_EX2>
__global__ void __launch_bounds__(64 + 32 * SW, 2)
k_mha_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
         int T, int heads, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = base;
    uint8_t* sK = sQ + kQBytes;
    uint8_t* sV = sK + kKBytes;
    uint8_t* sP = sV + kVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4,
             *s_full = bars + 5, *p_ready = bars + 6, *o_full = bars + 7;      // s_full[0], and s_full[1] = bars + 8
    uint64_t* s_full2[2] = {s_full, bars + 8};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* s_x = reinterpret_cast<float*>(bars + 10);       // [2][128] exchange of row maxima / sums between the two halves

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kMhaQ;
    const int hh = blockIdx.y, b = blockIdx.z;
    const int D = heads * kMhaD;
    const int ntiles = (T + kMhaK - 1) / kMhaK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full2[0], 1);
        mbar_init(s_full2[1], 1);
        mbar_init(p_ready, 32 * SW);
        mbar_init(o_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = *tmem_slot;                    // columns [0, 64) and [64, 128): S (double-buffered), [128, 192): O tile
    const uint32_t tmem_o = tmem_s + 128;

    // The two single-issuer roles run with the WHOLE warp converged and issue under elect.sync (tc_common.cuh elect_one): the
    // operands stay in uniform registers instead of an ELECT + BRA.U.ANY loop around every UTMALDG / UTCHMMA / UTCBAR.
    if (warp == 0) {                                       // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(q_full, kQBytes);
            tma_load_3d(sQ, &tmQ, q_full, hh * kMhaD, q0, b);
            tma_load_3d(sQ + kQBytes / 2, &tmQ, q_full, hh * kMhaD + 32, q0, b);
        }
        __syncwarp();
        for (int j = 0; j < ntiles; ++j) {
            const int k0 = j * kMhaK;
            if (j > 0) mbar_wait(k_empty, (j - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(k_full, kKBytes);
                tma_load_3d(sK, &tmK, k_full, D + hh * kMhaD, k0, b);
                tma_load_3d(sK + kKBytes / 2, &tmK, k_full, D + hh * kMhaD + 32, k0, b);
            }
            __syncwarp();
            if (j > 0) mbar_wait(v_empty, (j - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(v_full, kVBytes);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        tma_load_3d(sV + kb * 8192 + g * 4096, &tmV, v_full, 2 * D + hh * kMhaD + g * 32, k0 + kb * 32, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {                                // ===== MMA issuer =====
        constexpr uint32_t idesc_s = umma_idesc_tf32(kMhaQ, kMhaK);
        constexpr uint32_t idesc_o = umma_idesc_tf32_bmn(kMhaQ, kMhaD);
        mbar_wait(q_full, 0);
        // S of tile j + 1 is issued BEFORE the P V product of tile j (into the other S buffer), so it runs -- and its
        // completion reaches the softmax warps -- while they are still busy with tile j
        auto issue_s = [&](int jj) {
            mbar_wait(k_full, jj & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t da = umma_desc_k128(smem_u32(sQ + kb * (kQBytes / 2)));
                    const uint64_t db = umma_desc_k128(smem_u32(sK + kb * (kKBytes / 2)));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_tf32(tmem_s + (uint32_t)((jj & 1) * 64), da + 2 * k, db + 2 * k, idesc_s, (kb | k) != 0);
                }
                umma_commit(k_empty);
                umma_commit(s_full2[jj & 1]);
            }
            __syncwarp();
        };
        issue_s(0);
        for (int j = 0; j < ntiles; ++j) {
            if (j + 1 < ntiles) issue_s(j + 1);
            mbar_wait(p_ready, j & 1);
            mbar_wait(v_full, j & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t da = umma_desc_k128(smem_u32(sP + kb * (kPBytes / 2)));
                    const uint64_t db = umma_desc_mn128(smem_u32(sV + kb * 8192), 4096);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_tf32(tmem_o, da + 2 * k, db + 64 * k, idesc_o, (kb | k) != 0);
                }
                umma_commit(v_empty);
                umma_commit(o_full);                       // arrives last: the CTA outlives every pending arrive
            }
            __syncwarp();
        }
    } else {                                               // ===== softmax / output: warps 2 .. 2 + SW =====
        constexpr int HV = SW / 4;                         // threads per query row
        constexpr int NC = kMhaK / HV;                     // keys (and output columns) per thread: 64 or 32
        const int qd = warp & 3;                           // TMEM lane quarter this warp may access
        const int hf = (warp - 2) >> 2;                    // which half of the columns (0 when SW == 4)
        const int row = qd * 32 + lane;                    // query row of the tile = TMEM lane
        const int col0 = hf * NC;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const float kScale = 0.125f * 1.4426950408889634f; // head_dim^-0.5 * log2(e)
        float o[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) o[i] = 0.0f;
        float m = -INFINITY, l = 0.0f;
        uint8_t* prow = sP + row * 128 + (col0 / 32) * (kPBytes / 2);
        const int sw = row & 7;
        for (int j = 0; j < ntiles; ++j) {
            const int k0 = j * kMhaK + col0;               // first key of this thread's columns
            mbar_wait(s_full2[j & 1], (j >> 1) & 1);
            tc_fence_after();
            float sv[NC];
#pragma unroll
            for (int blk = 0; blk < NC / 32; ++blk)
                tmem_ld32(tmem_s + lane_off + (uint32_t)((j & 1) * 64 + col0 + 32 * blk), reinterpret_cast<float(&)[32]>(sv[32 * blk]));
            if (j * kMhaK + kMhaK > T) {                   // ragged last tile only: keys past T drop out of max and sum
#pragma unroll
                for (int i = 0; i < NC; ++i)
                    if (k0 + i >= T) sv[i] = -INFINITY;
            }
            float mx = sv[0];
#pragma unroll
            for (int i = 1; i < NC; ++i) mx = fmaxf(mx, sv[i]);
            if (HV == 2) {                                 // the other half of the row lives in another warp
                s_x[hf * 128 + row] = mx;
                asm volatile("bar.sync 1, %0;" ::"r"(32 * SW) : "memory");
                mx = fmaxf(s_x[row], s_x[128 + row]);
            }
            const float mn = fmaxf(m, mx * kScale);        // kScale > 0: scaling commutes with max; finite (key 0 of the tile is valid)
            const float alpha = ex2<FAST_EX2>(m - mn);      // first tile: exp2(-inf) = 0
            m = mn;
            float sum = 0.0f;
#pragma unroll
            for (int blk = 0; blk < NC / 32; ++blk) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {              // 16-byte chunks of the 128-byte line of this row in K block blk
                    float4 p4;
                    p4.x = round_up_tf32(ex2<FAST_EX2>(fmaf(sv[32 * blk + 4 * c + 0], kScale, -mn)));
                    p4.y = round_up_tf32(ex2<FAST_EX2>(fmaf(sv[32 * blk + 4 * c + 1], kScale, -mn)));
                    p4.z = round_up_tf32(ex2<FAST_EX2>(fmaf(sv[32 * blk + 4 * c + 2], kScale, -mn)));
                    p4.w = round_up_tf32(ex2<FAST_EX2>(fmaf(sv[32 * blk + 4 * c + 3], kScale, -mn)));
                    sum += (p4.x + p4.y) + (p4.z + p4.w);
                    *reinterpret_cast<float4*>(prow + blk * (kPBytes / 2) + ((c ^ sw) << 4)) = p4;
                }
            }
            l = l * alpha + sum;                           // partial sum over this thread's columns
            fence_proxy_async_smem();                      // generic-proxy writes of P -> visible to the tensor core
            tc_fence_before();                             // the TMEM reads of S are done before the next S product
            mbar_arrive(p_ready);
            if (!__all_sync(0xffffffffu, alpha == 1.0f)) { // the running maxima settle after a few tiles
#pragma unroll
                for (int i = 0; i < NC; ++i) o[i] *= alpha;
            }
            mbar_wait(o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int blk = 0; blk < NC / 32; ++blk) {
                float t32[32];
                tmem_ld32(tmem_o + lane_off + (uint32_t)(col0 + 32 * blk), t32);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[32 * blk + i] += t32[i];
            }
            tc_fence_before();
        }
        if (HV == 2) {                                     // total row sum = the two partial sums (same maxima history)
            asm volatile("bar.sync 1, %0;" ::"r"(32 * SW) : "memory");      // every read of the last maxima exchange is done
            s_x[hf * 128 + row] = l;
            asm volatile("bar.sync 1, %0;" ::"r"(32 * SW) : "memory");
            l = s_x[row] + s_x[128 + row];
        }
        const int q = q0 + row;
        if (q < T) {
            const float inv = 1.0f / l;
            float* dst = out + ((int64_t)b * T + q) * D + hh * kMhaD + col0;
#pragma unroll
            for (int i = 0; i < NC; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_s, 256);
}

}  // namespace tc
}  // namespace oess

using namespace oess;

OESS_API int oess_mha_fwd_tc(const float* qkv, int B, int T, int heads, float* out, oess_stream_t stream) {
    if (B < 0 || T <= 0 || heads <= 0 || heads > 65535 || B > 65535) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!qkv || !out || (((uintptr_t)qkv | (uintptr_t)out) & 15)) return OESS_E_ARG;
    const int D = heads * tc::kMhaD;
    CUtensorMap tmQ, tmK, tmV;
    const uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)T, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)3 * D * 4, (uint64_t)T * 3 * D * 4};
    const uint32_t bq[3] = {32, tc::kMhaQ, 1}, bk[3] = {32, tc::kMhaK, 1}, bv[3] = {32, 32, 1};
    int rc = tc::make_tmap_f32(&tmQ, qkv, 3, dims, strides, bq);
    if (rc) return rc;
    rc = tc::make_tmap_f32(&tmK, qkv, 3, dims, strides, bk);
    if (rc) return rc;
    rc = tc::make_tmap_f32_atom32(&tmV, qkv, 3, dims, strides, bv);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    static const int sw = (getenv("OESS_MHA_WARPS") && atoi(getenv("OESS_MHA_WARPS")) == 4) ? 4 : 8;
    const dim3 grid((unsigned)((T + tc::kMhaQ - 1) / tc::kMhaQ), (unsigned)heads, (unsigned)B);
    static const bool fast = !(getenv("OESS_MHA_EX2") && getenv("OESS_MHA_EX2")[0] == 'e');
#define OESS_MHA_LAUNCH(SWV, FV)                                                                                                   \
    do {                                                                                                                           \
        OESS_CUDA(cudaFuncSetAttribute(tc::k_mha_tc<SWV, FV>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kMhaSmemBytes));       \
        OESS_KERNEL("mha_fwd_tc", st, tc::k_mha_tc<SWV, FV><<<grid, 64 + 32 * SWV, tc::kMhaSmemBytes, st>>>(tmQ, tmK, tmV, T, heads, out)); \
    } while (0)
    if (sw == 8) {
        if (fast) OESS_MHA_LAUNCH(8, true); else OESS_MHA_LAUNCH(8, false);
    } else {
        if (fast) OESS_MHA_LAUNCH(4, true); else OESS_MHA_LAUNCH(4, false);
    }
#undef OESS_MHA_LAUNCH
    return OESS_OK;
}
