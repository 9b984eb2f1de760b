// tc_conv.cu -- 2-D convolution over channels-last activations as a tcgen05 implicit GEMM (TF32 operands, fp32 accumulate):
//     y[b, oy, ox, :] = act( bias + sum_{ky,kx,c} w[:, ky, kx, c] * x[b, oy*s - p + ky*d, ox*s - p + kx*d, c]  (+ residual) )
// Serves the frozen / inference convolutions of the path: the strided 5x5 encoder convolutions of E2VID with folded
// eval-mode BatchNorm + ReLU (e2vid/model/submodules.py:7-31, unet.py:128-135), and any 1x1 / 3x3 (dilated) convolution
// with Cin % 4 == 0 (models/_resnet.py bottlenecks, models/deeplabv3.py ASPP).
//
// Same skeleton as tc_convlstm.cu: M = 128 output pixels = an 8 x 16 patch, N = 64 / 128 / 256 output channels,
// K = taps x Cin (Cin padded to 32 per tap in the packed weights; the activation box is zero-filled beyond Cin by TMA).
// The A tile of one (tap, 32-channel chunk) is ONE 4-D TMA box {32 ch, 16 s, 8 s, 1} with element strides {1, s, s, 1}
// (stride-s convolution = strided box traversal) at coordinates shifted by the tap (dilation d): the TMA unit's
// out-of-bounds zero fill is the convolution's zero padding.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace oess {
namespace tc {

constexpr int kVW = 16, kVH = 8;
constexpr int kVABytes = 128 * kBlockK * 4;

// ring depth chosen so that TWO CTAs fit one SM (<= ~100 KB each): one CTA's epilogue overlaps the other's main loop
template <int BN>
struct ConvSmem {
    static constexpr int kBBytes = BN * kBlockK * 4;
    static constexpr int kStages = BN >= 256 ? 2 : (BN >= 128 ? 3 : (BN >= 64 ? 4 : 5));
    static constexpr int kStatBytes = 2 * 4 * BN * 4;      // per epilogue warp: column sums and sums of squares
    static constexpr int kBytes = 1024 + kStages * (kVABytes + kBBytes) + 256 + kStatBytes;
};

struct ConvArgs {
    int Ho, Wo, Cout, KW, taps, chunks, stride, pad, pad_x, dil, relu, tiles_w;   // pad: rows, pad_x: columns
    int stats_stride;   // doubles between the statistics of consecutive samples (0: one set for the whole batch)
};

// BF16: bfloat16 activations and packed weights (tcgen05.mma.kind::f16, 64 elements per 128-byte K block, fp32 accumulate) for
// the FROZEN networks (teacher ResNet-50, E2VID encoder convs): half the operand bytes and twice the MMA rate of TF32.
template <int BN, bool BF16>
__global__ void __launch_bounds__(192, 2)
k_conv_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
          const float* __restrict__ residual, float* __restrict__ y, __nv_bfloat16* __restrict__ ybf,
          double* __restrict__ bn_sums, const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    using S = ConvSmem<BN>;
    constexpr int kVStages = S::kStages;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kVStages * kVABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kVStages * S::kBBytes);
    uint64_t* empty = full + kVStages;
    uint64_t* acc_full = empty + kVStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* s_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);   // [2][4 warps][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int th = blockIdx.x / a.tiles_w, tw = blockIdx.x - th * a.tiles_w;
    const int h0 = th * kVH, w0 = tw * kVW;
    const int n0 = blockIdx.y * BN, b = blockIdx.z;
    const int kblocks = a.taps * a.chunks;
    constexpr int kKE = BF16 ? 64 : kBlockK;              // operand elements per 128-byte K block

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kVStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                  // ===== TMA producer =====
            int tap = 0, chunk = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kVStages;
                mbar_wait(&empty[s], ((kb / kVStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], kVABytes + S::kBBytes);
                const int ky = tap / a.KW, kx = tap - ky * a.KW;
                tma_load_4d(sA + s * kVABytes, &tmX, &full[s], chunk * kKE, w0 * a.stride - a.pad_x + kx * a.dil,
                            h0 * a.stride - a.pad + ky * a.dil, b);
                tma_load_2d(sB + s * S::kBBytes, &tmW, &full[s], kb * kKE, n0);
                if (++chunk == a.chunks) { chunk = 0; ++tap; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                  // ===== MMA issuer =====
            constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(128, BN) : umma_idesc_tf32(128, BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kVStages;
                mbar_wait(&full[s], (kb / kVStages) & 1);
                tc_fence_after();
                const uint64_t da = umma_desc_k128(smem_u32(sA + s * kVABytes));
                const uint64_t db = umma_desc_k128(smem_u32(sB + s * S::kBBytes));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {   // four 32-byte K steps per block: K = 8 (tf32) / 16 (bf16) each
                    if (BF16) umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    else umma_tf32(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(acc_full);
        }
    } else {                                              // ===== epilogue: warps 2..5 =====
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int r = q * 32 + lane;
        const int oy = h0 + r / kVW, ox = w0 + r % kVW;
        const bool valid = oy < a.Ho && ox < a.Wo;
        const int64_t pix = (((int64_t)b * a.Ho + oy) * a.Wo + ox) * a.Cout;
        const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
        const bool vec = (a.Cout & 3) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            tmem_ld16_nowait(trow + c0, v);
            tmem_ld_wait();
            const int col = n0 + c0;
            if (bn_sums) {
                // BatchNorm batch statistics of the raw conv output, fused: per-column sum / sum of squares over the
                // warp's 32 rows by recursive halving (16 shuffles each), staged per warp in shared memory.
                float s8[16], q8[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float t = valid ? v[j] + ((bias && col + j < a.Cout) ? __ldg(bias + col + j) : 0.f) : 0.f;   // statistics of y
                    s8[j] = t;
                    q8[j] = t * t;
                }
#pragma unroll
                for (int step = 0; step < 4; ++step) {                 // xor 16, 8, 4, 2: keep half of the columns
                    const int m = 16 >> step, half = 8 >> step;
                    const bool up = (lane & m) != 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < half) {
                            const float send_s = up ? s8[j] : s8[j + half], send_q = up ? q8[j] : q8[j + half];
                            const float keep_s = up ? s8[j + half] : s8[j], keep_q = up ? q8[j + half] : q8[j];
                            s8[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, m);
                            q8[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, m);
                        }
                    }
                }
                s8[0] += __shfl_xor_sync(0xffffffffu, s8[0], 1);
                q8[0] += __shfl_xor_sync(0xffffffffu, q8[0], 1);
                if ((lane & 1) == 0) {
                    const int cj = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    s_stat[q * BN + c0 + cj] = s8[0];
                    s_stat[(4 + q) * BN + c0 + cj] = q8[0];
                }
            }
            if (!valid || col >= a.Cout) continue;
            if (vec && col + 16 <= a.Cout) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col + j));
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (residual) {
                        const float4 rr = *reinterpret_cast<const float4*>(residual + pix + col + j);
                        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                    }
                    if (a.relu & 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (a.relu & 2) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
                    if (y) *reinterpret_cast<float4*>(y + pix + col + j) = o;
                    if (ybf) {                                        // bf16 copy: operand of a kind::f16 consumer
                        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(ybf + pix + col + j) = pk;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (col + j < a.Cout) {
                        float o = v[j] + (bias ? bias[col + j] : 0.f) + (residual ? residual[pix + col + j] : 0.f);
                        if (a.relu & 1) o = fmaxf(o, 0.f);
                        if (y) y[pix + col + j] = (a.relu & 2) ? rna_tf32(o) : o;
                        if (ybf) ybf[pix + col + j] = __float2bfloat16_rn(o);
                    }
                }
            }
        }
        if (bn_sums) {
            asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps only
            for (int c = (int)threadIdx.x - 64; c < BN; c += 128) {
                if (n0 + c < a.Cout) {
                    const float ts = s_stat[c] + s_stat[BN + c] + s_stat[2 * BN + c] + s_stat[3 * BN + c];
                    const float tq = s_stat[4 * BN + c] + s_stat[5 * BN + c] + s_stat[6 * BN + c] + s_stat[7 * BN + c];
                    double* sums = bn_sums + (size_t)b * a.stats_stride;
                    atomicAdd(&sums[n0 + c], (double)ts);
                    atomicAdd(&sums[a.Cout + n0 + c], (double)tq);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, BN);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default when Cout % 4 == 0; OESS_CONV=tile selects the kernel above): one CTA per SM walks the tiles
// (pixel tile fastest, then Cout tile, then sample), the operand ring keeps running across tile boundaries, two TMEM
// accumulators: epilogue group g (four warps, one per TMEM lane quarter) drains the tiles of accumulator g while the MMA warp
// fills the other one.  fp32 output goes through a swizzled 32-pixel x 32-channel staging block per warp and a 4-D TMA store
// (box {32 ch, 16 px, 2 rows, 1}, clipped at the ragged edges); BatchNorm / InstanceNorm statistics are reduced across the
// warp's 32 pixels by recursive halving (lane l ends up with channel l of the chunk), kept in registers across tiles and
// flushed with fp64 atomics only when the CTA moves to another Cout tile or sample.
constexpr int kPConvThreads = 320;

template <int BN, bool PAIR = false>
struct PConvSmem {
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBlockK * 4;   // CTA pair: this CTA's half of the weight tile
    // narrow tiles: TWO K blocks per ring stage.  Their four MMAs take 4 x BN / 2 cycles, far less than one trip of the
    // single-lane issue loops (~430 cycles: wait, elect, descriptors, commit), so the loops -- not the tensor pipe -- set the
    // pace; eight MMAs and four TMA boxes per trip halve that overhead.
    static constexpr int kKbPerStage = BN <= 128 ? 2 : 1;
    static constexpr int kStageBytes = kKbPerStage * (kVABytes + kBBytes);
    static constexpr int kStages = PAIR ? (BN >= 256 ? 6 : 4) : (BN >= 256 ? 4 : (BN >= 128 ? 3 : 4));
    static constexpr int kBytes = 1024 + kStages * kStageBytes + 8 * 4096 + 256;
};

struct PConvArgs {
    ConvArgs a;
    int tiles_px, n_tiles, tiles;
};

// PAIR (cta_group::2, default when there is enough work; OESS_CONV_2SM=0: off): the two CTAs of a cluster compute two neighbouring
// pixel tiles x the same Cout tile as ONE M = 256 MMA.  Each CTA stages its own A box and HALF of the weight tile, so the ring is
// 6 (BN = 256) / 4 x 2 (BN = 128) K blocks deep instead of 4 / 3 x 2 and every SM reads a third less shared memory per MMA; the
// leader issues the MMAs and commits to both CTAs' barriers, every CTA's epilogue drains its own 128 TMEM lanes.
template <int BN, bool BF16, bool PAIR>
__global__ void __launch_bounds__(kPConvThreads, 1)
k_conv_tc_p(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmY, const float* __restrict__ bias, const float* __restrict__ residual,
            const bool y_f32, __nv_bfloat16* __restrict__ ybf, double* __restrict__ bn_sums, const PConvArgs pa) {
    extern __shared__ uint8_t smem_raw[];
    using S = PConvSmem<BN, PAIR>;
    constexpr int kStages = S::kStages;
    constexpr int kKE = BF16 ? 64 : kBlockK;
    constexpr int NCH = BN / 32;
    const ConvArgs& a = pa.a;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    // PAIR: `tile` counts pixel-tile PAIRS (pixel tiles 2 pu + rank; one past the end = a dummy tile: zero-filled loads, clipped stores)
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int px_units = PAIR ? (pa.tiles_px + 1) / 2 : pa.tiles_px;
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int KPS = S::kKbPerStage;
    uint8_t* sA = base;                                   // [kStages][KPS] A boxes
    uint8_t* sB = sA + kStages * KPS * kVABytes;          // [kStages][KPS] B boxes
    uint8_t* sC = sB + kStages * KPS * S::kBBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + 8 * 4096);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = a.taps * a.chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        if (y_f32) tma_prefetch_desc(&tmY);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], PAIR ? 8 : 4);       // PAIR: both CTAs' epilogue warps arrive on the leader's barrier
        }
        mbar_fence_init();
    }
    if (PAIR) {
        cluster_sync_all();                               // barriers of both CTAs exist before TMEM allocation / any remote arrive
        if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN);
        tc_fence_before();
        cluster_sync_all();
    } else {
        if (warp == 1) tmem_alloc(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN);
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {                                      // ===== TMA producer (whole warp converged, one lane issues) =====
        uint32_t s = 0, ph = 1;                           // ring slot and the parity to wait for on its empty barrier
        for (int tile = tile0; tile < pa.tiles; tile += tstep) {
            const int px = PAIR ? 2 * (tile % px_units) + (int)crank : tile % px_units, rest = tile / px_units;
            const int n0 = (rest % pa.n_tiles) * BN + (PAIR ? (int)crank * (BN / 2) : 0), b = rest / pa.n_tiles;
            const int th = px / a.tiles_w, tw = px - th * a.tiles_w;
            const int x0 = tw * kVW * a.stride - a.pad_x, y0 = th * kVH * a.stride - a.pad;
            int ky = 0, kx = 0, chunk = 0;
            for (int kb = 0; kb < kblocks; kb += KPS) {
                mbar_wait(&empty[s], ph);
                const int nk = (KPS == 1 || kb + 1 < kblocks) ? KPS : 1;      // K blocks in this stage (odd tail: one)
                int cc[KPS], cx[KPS], cy[KPS];                                // the tap walk, by every lane alike
#pragma unroll
                for (int u = 0; u < KPS; ++u) {
                    cc[u] = chunk * kKE;
                    cx[u] = x0 + kx * a.dil;
                    cy[u] = y0 + ky * a.dil;
                    if (u < nk && ++chunk == a.chunks) { chunk = 0; if (++kx == a.KW) { kx = 0; ++ky; } }
                }
                if (elect_one()) {
                    if (!PAIR) mbar_expect_tx(&full[s], nk * (kVABytes + S::kBBytes));
                    else if (leader) mbar_expect_tx(&full[s], 2 * nk * (kVABytes + S::kBBytes));      // both CTAs' bytes
#pragma unroll
                    for (int u = 0; u < KPS; ++u) {
                        if (u < nk) {
                            if (PAIR) {
                                tma_load_4d_2sm(sA + (s * KPS + u) * kVABytes, &tmX, &full[s], cc[u], cx[u], cy[u], b);
                                tma_load_2d_2sm(sB + (s * KPS + u) * S::kBBytes, &tmW, &full[s], (kb + u) * kKE, n0);
                            } else {
                                tma_load_4d(sA + (s * KPS + u) * kVABytes, &tmX, &full[s], cc[u], cx[u], cy[u], b);
                                tma_load_2d(sB + (s * KPS + u) * S::kBBytes, &tmW, &full[s], (kb + u) * kKE, n0);
                            }
                        }
                    }
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {                               // ===== MMA issuer (whole warp converged, one lane issues; PAIR: leader only) =====
        constexpr int kMmaM = PAIR ? 256 : 128;
        constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(kMmaM, BN) : umma_idesc_tf32(kMmaM, BN);
        uint32_t s = 0, ph = 0, lt = 0;
        for (int tile = tile0; leader && tile < pa.tiles; tile += tstep, ++lt) {
            const uint32_t buf = lt & 1;
            mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_acc + buf * BN;
            for (int kb = 0; kb < kblocks; kb += KPS) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const int nk = (KPS == 1 || kb + 1 < kblocks) ? KPS : 1;
                if (elect_one()) {
#pragma unroll
                    for (int u = 0; u < KPS; ++u) {
                        if (u < nk) {
                            const uint64_t da = umma_desc_k128(smem_u32(sA + (s * KPS + u) * kVABytes));
                            const uint64_t db = umma_desc_k128(smem_u32(sB + (s * KPS + u) * S::kBBytes));
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                if (PAIR) {
                                    if (BF16) umma_bf16_2sm(d, da + 2 * k, db + 2 * k, idesc, (kb | u | k) != 0);
                                    else umma_tf32_2sm(d, da + 2 * k, db + 2 * k, idesc, (kb | u | k) != 0);
                                } else {
                                    if (BF16) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb | u | k) != 0);
                                    else umma_tf32(d, da + 2 * k, db + 2 * k, idesc, (kb | u | k) != 0);
                                }
                            }
                        }
                    }
                    if (PAIR) umma_commit_2sm(&empty[s], (uint16_t)3);
                    else umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) {
                if (PAIR) umma_commit_2sm(&acc_full[buf], (uint16_t)3);
                else umma_commit(&acc_full[buf]);
            }
            __syncwarp();
        }
    } else {                                              // ===== epilogue: group g = warps 2 + 4 g .. 5 + 4 g =====
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        const uint32_t g = (uint32_t)(warp - 2) >> 2;     // drains accumulator g = the CTA's tiles with (local index & 1) == g
        uint8_t* stage = sC + (warp - 2) * 4096;
        float st_s[NCH], st_q[NCH];                       // running statistics of channel n0 + 32 ch + lane over this warp's pixels
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) st_s[ch] = st_q[ch] = 0.f;
        int st_key = -1;                                  // (sample, Cout tile) the running statistics belong to
        auto flush = [&](int key) {
            if (key < 0) return;
            const int n0 = (key % pa.n_tiles) * BN, b = key / pa.n_tiles;
            double* sums = bn_sums + (size_t)b * a.stats_stride;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = n0 + ch * 32 + lane;
                if (c < a.Cout) {
                    atomicAdd(&sums[c], (double)st_s[ch]);
                    atomicAdd(&sums[a.Cout + c], (double)st_q[ch]);
                }
                st_s[ch] = st_q[ch] = 0.f;
            }
        };
        uint32_t lt = 0;
        for (int tile = tile0; tile < pa.tiles; tile += tstep, ++lt) {
            if ((lt & 1) != g) continue;
            const int px = PAIR ? 2 * (tile % px_units) + (int)crank : tile % px_units, rest = tile / px_units;
            const int n0 = (rest % pa.n_tiles) * BN, b = rest / pa.n_tiles;
            const int th = px / a.tiles_w, tw = px - th * a.tiles_w;
            const int h0 = th * kVH, w0 = tw * kVW;
            if (bn_sums && rest != st_key) {
                flush(st_key);
                st_key = rest;
            }
            mbar_wait(&acc_full[g], (lt >> 1) & 1);
            tc_fence_after();
            const int oy = h0 + q * 2 + (lane >> 4), ox = w0 + (lane & 15);
            const bool valid = oy < a.Ho && ox < a.Wo;
            const int64_t pix = (((int64_t)b * a.Ho + oy) * a.Wo + ox) * a.Cout;
            const uint32_t t0 = tmem_acc + g * BN + ((uint32_t)(q * 32) << 16);
            float v[2][32];
            tmem_ld32_nowait(t0, v[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                tmem_ld_wait();
                if (ch + 1 < NCH) tmem_ld32_nowait(t0 + (uint32_t)(ch + 1) * 32, v[(ch + 1) & 1]);
                if (ch + 1 == NCH) {                      // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(&acc_empty[g], 0);
                        else mbar_arrive(&acc_empty[g]);
                    }
                }
                const int col = n0 + ch * 32;
                if (col >= a.Cout) continue;              // warp-uniform
                float (&w)[32] = v[ch & 1];
                const bool full32 = col + 32 <= a.Cout;
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (full32 || col + j + 4 <= a.Cout) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col + j));
                            w[j] += bb.x; w[j + 1] += bb.y; w[j + 2] += bb.z; w[j + 3] += bb.w;
                        }
                    }
                }
                if (bn_sums) {
                    // column sums / sums of squares over the warp's 32 pixels: five exchange rounds, each halving the columns
                    // a lane is responsible for; lane l ends with channel col + l
                    float s_[16], q_[16];
                    {
                        const bool up = (lane & 16) != 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float lo = valid ? w[j] : 0.f, hi = valid ? w[j + 16] : 0.f;
                            const float keep = up ? hi : lo, send = up ? lo : hi;
                            s_[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                            q_[j] = keep * keep + __shfl_xor_sync(0xffffffffu, send * send, 16);
                        }
                    }
#pragma unroll
                    for (int step = 1; step < 5; ++step) {
                        const int m = 16 >> step, hcount = 16 >> step;
                        const bool up = (lane & m) != 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j < hcount) {
                                const float ks = up ? s_[j + hcount] : s_[j], ss = up ? s_[j] : s_[j + hcount];
                                const float kq = up ? q_[j + hcount] : q_[j], sq = up ? q_[j] : q_[j + hcount];
                                s_[j] = ks + __shfl_xor_sync(0xffffffffu, ss, m);
                                q_[j] = kq + __shfl_xor_sync(0xffffffffu, sq, m);
                            }
                        }
                    }
                    st_s[ch] += s_[0];
                    st_q[ch] += q_[0];
                }
                if (residual && valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (full32 || col + j + 4 <= a.Cout) {
                            const float4 rr = *reinterpret_cast<const float4*>(residual + pix + col + j);
                            w[j] += rr.x; w[j + 1] += rr.y; w[j + 2] += rr.z; w[j + 3] += rr.w;
                        }
                    }
                }
                if (a.relu & 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[j] = fmaxf(w[j], 0.f);
                }
                if (a.relu & 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) w[j] = rna_tf32(w[j]);
                }
                if (ybf && valid) {                       // bf16 copy: operand of a kind::f16 consumer
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (full32 || col + j + 4 <= a.Cout) {
                            const __nv_bfloat162 lo = __floats2bfloat162_rn(w[j], w[j + 1]), hi = __floats2bfloat162_rn(w[j + 2], w[j + 3]);
                            uint2 pk;
                            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(ybf + pix + col + j) = pk;
                        }
                    }
                }
                if (y_f32) {
                    if (elect_one()) tma_store_wait_read<0>();   // this warp's previous store has read the staging block
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)       // 128-byte swizzle: 16-byte chunk j / 4 of row `lane` -> chunk (j / 4) ^ (lane & 7)
                        *reinterpret_cast<float4*>(stage + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) =
                            make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (elect_one()) {                    // always the same lane: bulk async-groups are per thread
                        tma_store_4d(&tmY, stage, col, w0, h0 + q * 2, b);
                        tma_store_commit();
                    }
                    __syncwarp();
                }
            }
        }
        if (bn_sums) flush(st_key);
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }
    tc_fence_before();
    if (PAIR) {
        cluster_sync_all();                               // no CTA leaves (or frees TMEM) while its peer still works
        if (warp == 1) tmem_dealloc_2sm(tmem_acc, 2 * BN < 32 ? 32 : 2 * BN);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_acc, 2 * BN < 32 ? 32 : 2 * BN);
    }
}

template <int BN, bool BF16 = false>
static int launch_conv(const CUtensorMap& tmX, const void* w_packed, int Cout, int Ktot, const float* bias,
                       const float* residual, float* y, double* bn_sums, const ConvArgs& a, int B, cudaStream_t st,
                       __nv_bfloat16* ybf = nullptr) {
    CUtensorMap tmW;
    const uint64_t dW[2] = {(uint64_t)Ktot, (uint64_t)Cout}, sW[1] = {(uint64_t)Ktot * (BF16 ? 2 : 4)};
    const uint32_t bW[2] = {BF16 ? 64u : (uint32_t)kBlockK, (uint32_t)BN};
    int rc = BF16 ? make_tmap_bf16(&tmW, w_packed, 2, dW, sW, bW) : make_tmap_f32(&tmW, w_packed, 2, dW, sW, bW);
    if (rc) return rc;
    const int tiles_h = (a.Ho + kVH - 1) / kVH;
    const int n_tiles = (Cout + BN - 1) / BN;
    static const bool tile_env = [] { const char* e = std::getenv("OESS_CONV"); return e && e[0] == 't'; }();
    const int64_t tiles = (int64_t)a.tiles_w * tiles_h * n_tiles * B;
    if (!tile_env && (Cout & 3) == 0 && tiles < (1ll << 31)) {
        CUtensorMap tmY;
        memset(&tmY, 0, sizeof(tmY));
        if (y) {
            const uint64_t dY[4] = {(uint64_t)Cout, (uint64_t)a.Wo, (uint64_t)a.Ho, (uint64_t)B};
            const uint64_t sY[3] = {(uint64_t)Cout * 4, (uint64_t)a.Wo * Cout * 4, (uint64_t)a.Ho * a.Wo * Cout * 4};
            const uint32_t bY[4] = {32, (uint32_t)kVW, 2, 1};
            rc = make_tmap_f32(&tmY, y, 4, dY, sY, bY);
            if (rc) return rc;
        }
        // CTA pairs: 256-wide tiles with K >= 1024 and at least one pair tile per pair of SMs.  Measured (profiles/r02_conv_pair.log):
        // 2048 -> 512 1 x 1 +14 %, 512 -> 512 3 x 3 +5 %, 1024 -> 256 +4 %; narrower tiles and K <= 512 lose 2 - 13 % (half as many
        // schedulable units, cluster-wide barriers per K block), so they stay on single CTAs.
        static const bool sm2_env = [] { const char* e = std::getenv("OESS_CONV_2SM"); return !e || e[0] != '0'; }();
        const int tiles_px = a.tiles_w * tiles_h;
        const int64_t pair_tiles = (int64_t)((tiles_px + 1) / 2) * n_tiles * B;
        if constexpr (BN == 256) if (sm2_env && a.taps * a.chunks * (BF16 ? 64 : kBlockK) >= 1024 && pair_tiles >= kNumSMs / 2) {
            CUtensorMap tmW2;                             // this CTA's half of the weight tile per box
            const uint32_t bW2[2] = {BF16 ? 64u : (uint32_t)kBlockK, (uint32_t)(BN / 2)};
            rc = BF16 ? make_tmap_bf16(&tmW2, w_packed, 2, dW, sW, bW2) : make_tmap_f32(&tmW2, w_packed, 2, dW, sW, bW2);
            if (rc) return rc;
            auto kern2 = k_conv_tc_p<BN, BF16, true>;
            OESS_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, PConvSmem<BN, true>::kBytes));
            const PConvArgs pa2{a, tiles_px, n_tiles, (int)pair_tiles};
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(kNumSMs & ~1));
            cfg.blockDim = dim3(kPConvThreads);
            cfg.dynamicSmemBytes = PConvSmem<BN, true>::kBytes;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const bool y_f32 = y != nullptr;
            OESS_KERNEL(BF16 ? "tc_conv2d_bf16" : "tc_conv2d", st,
                        cudaLaunchKernelEx(&cfg, kern2, tmX, tmW2, tmY, bias, residual, y_f32, ybf, bn_sums, pa2));
            return 0;
        }
        auto kern = k_conv_tc_p<BN, BF16, false>;
        OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PConvSmem<BN>::kBytes));
        const PConvArgs pa{a, a.tiles_w * tiles_h, n_tiles, (int)tiles};
        const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
        OESS_KERNEL(BF16 ? "tc_conv2d_bf16" : "tc_conv2d", st,
                    kern<<<grid, kPConvThreads, PConvSmem<BN>::kBytes, st>>>(tmX, tmW, tmY, bias, residual, y != nullptr, ybf,
                                                                             bn_sums, pa));
        return 0;
    }
    auto kern = k_conv_tc<BN, BF16>;
    OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<BN>::kBytes));
    const dim3 grid((unsigned)(a.tiles_w * tiles_h), (unsigned)n_tiles, (unsigned)B);
    OESS_KERNEL(BF16 ? "tc_conv2d_bf16" : "tc_conv2d", st,
                kern<<<grid, 192, ConvSmem<BN>::kBytes, st>>>(tmX, tmW, bias, residual, y, ybf, bn_sums, a));
    return 0;
}

// strided variant of make_tmap_f32 (element strides per dimension)
static int make_tmap_f32_strided(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box, const uint32_t* estr, bool bf16 = false) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = estr[i];
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                           const_cast<void*>(base), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace oess

using namespace oess;

// x: [B, H, W, Cin] channels-last; w_packed: [Cout, KH * KW * Cin_p] (Cin_p = Cin rounded up to 32, zero padded; column
// (tap = ky * KW + kx, channel)); bias [Cout] or NULL; residual [B, Ho, Wo, Cout] or NULL; y: [B, Ho, Wo, Cout].
static int conv2d_impl(const float* x, const float* w_packed, const float* bias, const float* residual, float* y,
                       int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                       int relu, double* bn_sums, int per_sample, oess_stream_t stream, __nv_bfloat16* ybf = nullptr) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || dil <= 0 || pad < 0)
        return OESS_E_ARG;
    if (!x || !w_packed || (!y && !ybf)) return OESS_E_ARG;
    if ((Cin & 3) || stride > 8 || KH * KW > 64) return OESS_E_ARG;      // TMA: 16-byte pixel stride; box <= 256 per dim
    if (((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)y | (uintptr_t)ybf) & 15) return OESS_E_ARG;
    if (B > 65535) return OESS_E_RANGE;
    const int Ho = (H + 2 * pad - dil * (KH - 1) - 1) / stride + 1;
    const int Wo = (W + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (Cin + tc::kBlockK - 1) / tc::kBlockK;
    const int Ktot = KH * KW * chunks * tc::kBlockK;
    CUtensorMap tmX;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    const uint32_t box[4] = {tc::kBlockK, (uint32_t)(tc::kVW * stride), (uint32_t)(tc::kVH * stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = tc::make_tmap_f32_strided(&tmX, x, 4, dims, strides, box, estr);
    if (rc) return rc;
    tc::ConvArgs a{Ho, Wo, Cout, KW, KH * KW, chunks, stride, pad, pad, dil, relu & 3, (Wo + tc::kVW - 1) / tc::kVW,
                   per_sample ? 2 * Cout : 0};
    if (bn_sums) OESS_CUDA(cudaMemsetAsync(bn_sums, 0, sizeof(double) * 2 * (size_t)Cout * (per_sample ? B : 1), st));
    if (Cout > 128) return tc::launch_conv<256>(tmX, w_packed, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    if (Cout > 64) return tc::launch_conv<128>(tmX, w_packed, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    if (Cout > 32) return tc::launch_conv<64>(tmX, w_packed, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    return tc::launch_conv<32>(tmX, w_packed, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
}

// Same convolution (TF32 operands) whose output is ALSO / ONLY stored as bfloat16 (y may be NULL): the operand of a bf16
// tensor-core consumer (oess_convlstm_step_nhwc_bf16).  Cout % 4 == 0 for the vector stores.
OESS_API int oess_conv2d_nhwc_tf32_bf16out(const float* x, const float* w_packed, const float* bias, float* y, void* y_bf16,
                                           int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                                           int relu, oess_stream_t stream) {
    if (!y_bf16 || (Cout & 3)) return OESS_E_ARG;
    return conv2d_impl(x, w_packed, bias, nullptr, y, B, H, W, Cin, Cout, KH, KW, stride, pad, dil, relu, nullptr, 0, stream,
                       (__nv_bfloat16*)y_bf16);
}

// Thin-input convolution (E2VID head: 5 event channels padded to 8, 5 x 5 kernel, unet.py:126-127).  With one K block of 32
// channels per tap, 25 taps x 32 = 800 K columns carry 25 x 5 useful ones.  Here the KW taps of a kernel row are folded into the
// channel dimension instead: the A operand of tap ky is the OVERLAPPING window x[b, oy + ky - pad, ox .. ox + KW), all Cin
// channels = KW * Cin contiguous floats of the channels-last row -- a tensor map whose pixel stride (Cin floats) is smaller than
// its innermost extent (KW * Cin floats).  K = KH * roundup(KW * Cin, 32) (5 x 64 = 320 for the head).  x must be zero-padded
// by (KW - 1) / 2 pixels at both row ends: x [B, H, W + KW - 1, Cin] (oess_planes_to_nhwc_padded_w); rows are padded by the
// TMA out-of-bounds fill as usual.  w_packed: [Cout, KH * chunks * 32], column (ky, kx * Cin + c).  Stride 1, dilation 1.
OESS_API int oess_conv2d_nhwc_tf32_rowunfold(const float* x, const float* w_packed, const float* bias, float* y, int B, int H,
                                             int W, int Cin, int Cout, int KH, int KW, int relu, oess_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || !(KH & 1) || !(KW & 1)) return OESS_E_ARG;
    if (!x || !w_packed || !y || (Cin & 3) || KW * Cin > 256 || KH > 64) return OESS_E_ARG;
    if (((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)bias | (uintptr_t)y) & 15) return OESS_E_ARG;
    if (B > 65535) return OESS_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    const int Wp = W + KW - 1, Cv = KW * Cin;
    const int chunks = (Cv + tc::kBlockK - 1) / tc::kBlockK;
    const int Ktot = KH * chunks * tc::kBlockK;
    CUtensorMap tmX;
    const uint64_t dims[4] = {(uint64_t)Cv, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)Wp * Cin * 4, (uint64_t)H * Wp * Cin * 4};
    const uint32_t box[4] = {tc::kBlockK, (uint32_t)tc::kVW, (uint32_t)tc::kVH, 1};
    const uint32_t estr[4] = {1, 1, 1, 1};
    int rc = tc::make_tmap_f32_strided(&tmX, x, 4, dims, strides, box, estr);
    if (rc) return rc;
    tc::ConvArgs a{H, W, Cout, 1, KH, chunks, 1, (KH - 1) / 2, 0, 1, relu & 3, (W + tc::kVW - 1) / tc::kVW, 0};
    if (Cout > 128) return tc::launch_conv<256>(tmX, w_packed, Cout, Ktot, bias, nullptr, y, nullptr, a, B, st);
    if (Cout > 64) return tc::launch_conv<128>(tmX, w_packed, Cout, Ktot, bias, nullptr, y, nullptr, a, B, st);
    if (Cout > 32) return tc::launch_conv<64>(tmX, w_packed, Cout, Ktot, bias, nullptr, y, nullptr, a, B, st);
    return tc::launch_conv<32>(tmX, w_packed, Cout, Ktot, bias, nullptr, y, nullptr, a, B, st);
}

OESS_API int oess_conv2d_nhwc_tf32(const float* x, const float* w_packed, const float* bias, const float* residual, float* y,
                                   int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                                   int relu, oess_stream_t stream) {
    return conv2d_impl(x, w_packed, bias, residual, y, B, H, W, Cin, Cout, KH, KW, stride, pad, dil, relu, nullptr, 0, stream);
}

// Same, additionally accumulating the BatchNorm batch statistics of y (the RAW conv output: pass residual = NULL,
// relu = 0) into bn_sums[0..Cout) = sum_rows y, bn_sums[Cout..2 Cout) = sum_rows y^2 (zeroed here): the statistics pass
// of oess_batchnorm_nhwc fused into the conv epilogue (feed the result to oess_batchnorm_nhwc_sums).
OESS_API int oess_conv2d_nhwc_tf32_stats(const float* x, const float* w_packed, const float* bias, float* y, int B, int H,
                                         int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                                         double* bn_sums, oess_stream_t stream) {
    if (!bn_sums) return OESS_E_ARG;
    return conv2d_impl(x, w_packed, bias, nullptr, y, B, H, W, Cin, Cout, KH, KW, stride, pad, dil, 0, bn_sums, 0, stream);
}

// Same with PER-SAMPLE statistics, in_sums[b][0..Cout) / [b][Cout..2 Cout): the input of InstanceNorm2d
// (oess_instancenorm_nhwc_sums; models/style_networks.py:252-289 ReLUINSConv2d / INSResBlock).
OESS_API int oess_conv2d_nhwc_tf32_instats(const float* x, const float* w_packed, const float* bias, float* y, int B, int H,
                                           int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                                           double* in_sums, oess_stream_t stream) {
    if (!in_sums) return OESS_E_ARG;
    return conv2d_impl(x, w_packed, bias, nullptr, y, B, H, W, Cin, Cout, KH, KW, stride, pad, dil, 0, in_sums, 1, stream);
}

// bfloat16-operand convolution for the frozen networks: x [B, H, W, Cin] bf16 channels-last (Cin % 8 == 0), w_packed
// [Cout, KH * KW * Cin_p] bf16 (Cin_p = Cin rounded up to 64, zero padded), fp32 accumulate / bias / residual; the result goes
// to y (fp32, may be NULL) and / or y_bf16 (may be NULL; Cout % 4 == 0), optionally with the BatchNorm batch statistics of the
// raw output in bn_sums (then residual = NULL, relu = 0, as oess_conv2d_nhwc_tf32_stats).
OESS_API int oess_conv2d_nhwc_bf16(const void* x_bf16, const void* w_packed_bf16, const float* bias, const float* residual,
                                   float* y, void* y_bf16, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                   int pad, int dil, int relu, double* bn_sums, oess_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || dil <= 0 || pad < 0)
        return OESS_E_ARG;
    if (!x_bf16 || !w_packed_bf16 || (!y && !y_bf16)) return OESS_E_ARG;
    if ((Cin & 7) || stride > 8 || KH * KW > 64 || (y_bf16 && (Cout & 3))) return OESS_E_ARG;
    if (bn_sums && (residual || relu)) return OESS_E_ARG;
    if (((uintptr_t)x_bf16 | (uintptr_t)w_packed_bf16 | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)y | (uintptr_t)y_bf16) & 15)
        return OESS_E_ARG;
    if (B > 65535) return OESS_E_RANGE;
    const int Ho = (H + 2 * pad - dil * (KH - 1) - 1) / stride + 1;
    const int Wo = (W + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (Cin + 63) / 64;
    const int Ktot = KH * KW * chunks * 64;
    CUtensorMap tmX;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)(tc::kVW * stride), (uint32_t)(tc::kVH * stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = tc::make_tmap_f32_strided(&tmX, x_bf16, 4, dims, strides, box, estr, true);
    if (rc) return rc;
    tc::ConvArgs a{Ho, Wo, Cout, KW, KH * KW, chunks, stride, pad, pad, dil, relu & 1, (Wo + tc::kVW - 1) / tc::kVW, 0};
    if (bn_sums) OESS_CUDA(cudaMemsetAsync(bn_sums, 0, sizeof(double) * 2 * (size_t)Cout, st));
    __nv_bfloat16* ybf = (__nv_bfloat16*)y_bf16;
    if (Cout > 128) return tc::launch_conv<256, true>(tmX, w_packed_bf16, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    if (Cout > 64) return tc::launch_conv<128, true>(tmX, w_packed_bf16, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    if (Cout > 32) return tc::launch_conv<64, true>(tmX, w_packed_bf16, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
    return tc::launch_conv<32, true>(tmX, w_packed_bf16, Cout, Ktot, bias, residual, y, bn_sums, a, B, st, ybf);
}
