// tc_convlstm.cu -- one ConvLSTM step of the E2VID recurrent encoder as ONE tensor-core kernel:
//     gates = Conv3x3(cat(x, h_prev)) + bias;  c = sigmoid(f) * c_prev + sigmoid(i) * tanh(g);  h = sigmoid(o) * tanh(c)
// Replaces e2vid/model/submodules.py:175-214 (ConvLSTM.forward: torch.cat + cuDNN conv + ~10 pointwise kernels; the
// [B, 4C, H, W] gate tensor is never materialised here).  The three encoder levels run it 20 times per sample: it is
// 65 % of the encoder's FLOPs.
//
// Implicit GEMM on tcgen05, channels-last activations:
//   M = 128 pixels = an 8 x 16 spatial patch, N = 256 = 4 gates x 64 hidden channels, K = 9 taps x 2C input channels.
//   A tile of K block (source, tap, 32-channel chunk) = the patch shifted by the tap, fetched by ONE 4-D TMA box
//   {32 ch, 16 w, 8 h, 1 b} -- out-of-image coordinates are zero-filled by TMA, which IS the conv's zero padding.
//   B tile = 256 rows of the gate weights repacked once on the host to [4C (chunk, gate, c), (source, tap, ch)].
//   warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma.kind::tf32, fp32 accumulator in 256 TMEM columns),
//   warps 2..5 = epilogue: tcgen05.ld the four gates of 16 channels of one pixel, LSTM pointwise math, write h and c.
// fp32 operands are read as TF32 (torch's default cuDNN conv arithmetic on the reference's GPU run), fp32 accumulate.
#include <cuda_bf16.h>

#include <cstdlib>

#include "tc_common.cuh"

namespace oess {
namespace tc {

constexpr int kTW = 16, kTH = 8;            // spatial patch of one M tile (16 x 8 = 128 pixels)
constexpr int kCN = 256;                    // N tile: 4 gates x 64 hidden channels
constexpr int kCStages = 2;               // 2 x 48 KB per CTA -> two CTAs per SM: one's epilogue overlaps the other's main loop
constexpr int kCABytes = 128 * kBlockK * 4; // 16 KB
constexpr int kCBBytes = kCN * kBlockK * 4; // 32 KB
constexpr int kCSmem = 1024 + kCStages * (kCABytes + kCBBytes) + 256;

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }
// Gate activations of the bf16 variant: one MUFU.TANH each (tanh.approx.f32, max relative error 2^-11 -- below the bf16
// rounding of the operands that produced the gates); sigmoid(x) = 0.5 tanh(x / 2) + 0.5.  The exact forms (ex2 + rcp, tanhf's
// ~25-instruction sequence) made the epilogue ~120 instructions per hidden channel: at the C = 64 level, where a tile has only
// 18 K blocks of MMA work, the eight epilogue warps were busy 90 % of the time at 0.28 IPC and the MMA warp waited for a free
// accumulator (profiles/r02_convlstm_full.md).
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// TF32 variant: tanh(x) = 1 - 2 / (1 + e^(2x)) on ex2.approx + rcp (absolute error ~2e-7, saturates correctly at +-inf) instead
// of tanhf's branchy ~25-instruction sequence.
template <bool FAST> __device__ __forceinline__ float act_tanh(float x) {
    return FAST ? tanh_fast(x) : 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}
template <bool FAST> __device__ __forceinline__ float act_sigm(float x) {
    return FAST ? fmaf(0.5f, tanh_fast(0.5f * x), 0.5f) : sigm(x);
}

// BF16 = false: fp32 operands read as TF32 (32 channels per 128-byte K block).  BF16 = true: bf16 operands
// (tcgen05.mma.kind::f16, 64 channels per K block: half the operand bytes per flop, twice the MMA rate), fp32 accumulate,
// fp32 cell state; the hidden state is written as fp32 (h_out, may be NULL) and as bf16 (h_bf, the next step's operand).
template <bool BF16>
__global__ void __launch_bounds__(192, 2)
k_convlstm_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
              const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias, const float* __restrict__ c_prev,
              float* __restrict__ h_out, __nv_bfloat16* __restrict__ h_bf, float* __restrict__ c_out, int H, int W, int C,
              int has_h, int tiles_w) {
    constexpr int kKE = BF16 ? 64 : kBlockK;           // elements (channels) per K block
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kCStages * kCABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kCStages * kCBBytes);
    uint64_t* empty = full + kCStages;
    uint64_t* acc_full = empty + kCStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int th = blockIdx.x / tiles_w, tw = blockIdx.x - th * tiles_w;
    const int h0 = th * kTH, w0 = tw * kTW;
    const int nchunk = blockIdx.y, b = blockIdx.z;
    const int chunks = C / kKE;
    const int kblocks = (has_h ? 2 : 1) * 9 * chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kCStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kCN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                  // ===== TMA producer =====
            int src = 0, tap = 0, chunk = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kCStages;
                mbar_wait(&empty[s], ((kb / kCStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], kCABytes + kCBBytes);
                const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                tma_load_4d(sA + s * kCABytes, src ? &tmH : &tmX, &full[s], chunk * kKE, w0 + dx, h0 + dy, b);
                tma_load_2d(sB + s * kCBBytes, &tmW, &full[s], kb * kKE, nchunk * kCN);
                if (++chunk == chunks) { chunk = 0; if (++tap == 9) { tap = 0; ++src; } }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                  // ===== MMA issuer =====
            constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(128, kCN) : umma_idesc_tf32(128, kCN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kCStages;
                mbar_wait(&full[s], (kb / kCStages) & 1);
                tc_fence_after();
                const uint64_t da = umma_desc_k128(smem_u32(sA + s * kCABytes));
                const uint64_t db = umma_desc_k128(smem_u32(sB + s * kCBBytes));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {       // 4 instructions of 32 operand bytes per 128-byte line
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
        const int r = q * 32 + lane;                      // accumulator row = pixel of the patch (h-major, w-minor)
        const int y = h0 + r / kTW, x = w0 + r % kTW;
        const bool valid = y < H && x < W;
        const int64_t pix = (((int64_t)b * H + y) * W + x) * C + nchunk * 64;
        const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
        const float* bn = bias + nchunk * kCN;
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {               // 16 hidden channels at a time
            float gi[16], gf[16], go[16], gg[16];
            tmem_ld16_nowait(trow + 0 * 64 + sub * 16, gi);   // submodules.py:203 chunk order: in, remember, out, cell
            tmem_ld16_nowait(trow + 1 * 64 + sub * 16, gf);
            tmem_ld16_nowait(trow + 2 * 64 + sub * 16, go);
            tmem_ld16_nowait(trow + 3 * 64 + sub * 16, gg);
            tmem_ld_wait();
            if (valid) {
                const int64_t e = pix + sub * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c_prev) pc = *reinterpret_cast<const float4*>(c_prev + e + j);
                    const float4 bi = __ldg(reinterpret_cast<const float4*>(bn + 0 * 64 + sub * 16 + j));
                    const float4 bf = __ldg(reinterpret_cast<const float4*>(bn + 1 * 64 + sub * 16 + j));
                    const float4 bo = __ldg(reinterpret_cast<const float4*>(bn + 2 * 64 + sub * 16 + j));
                    const float4 bg = __ldg(reinterpret_cast<const float4*>(bn + 3 * 64 + sub * 16 + j));
                    float4 c, h;
                    c.x = act_sigm<BF16>(gf[j] + bf.x) * pc.x + act_sigm<BF16>(gi[j] + bi.x) * act_tanh<BF16>(gg[j] + bg.x);
                    c.y = act_sigm<BF16>(gf[j + 1] + bf.y) * pc.y + act_sigm<BF16>(gi[j + 1] + bi.y) * act_tanh<BF16>(gg[j + 1] + bg.y);
                    c.z = act_sigm<BF16>(gf[j + 2] + bf.z) * pc.z + act_sigm<BF16>(gi[j + 2] + bi.z) * act_tanh<BF16>(gg[j + 2] + bg.z);
                    c.w = act_sigm<BF16>(gf[j + 3] + bf.w) * pc.w + act_sigm<BF16>(gi[j + 3] + bi.w) * act_tanh<BF16>(gg[j + 3] + bg.w);
                    h.x = act_sigm<BF16>(go[j] + bo.x) * act_tanh<BF16>(c.x);
                    h.y = act_sigm<BF16>(go[j + 1] + bo.y) * act_tanh<BF16>(c.y);
                    h.z = act_sigm<BF16>(go[j + 2] + bo.z) * act_tanh<BF16>(c.z);
                    h.w = act_sigm<BF16>(go[j + 3] + bo.w) * act_tanh<BF16>(c.w);
                    *reinterpret_cast<float4*>(c_out + e + j) = c;
                    if (h_out) *reinterpret_cast<float4*>(h_out + e + j) = h;
                    if (BF16) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(h_bf + e + j) = pk;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, kCN);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default; OESS_CONVLSTM=tile selects the kernel above): one CTA per SM walks the tiles (pixel patch
// fastest, then 64-channel chunk, then sample) with a 4-stage operand ring that keeps running across tile boundaries and TWO
// TMEM accumulators (2 x 256 columns): epilogue group g (four warps, one per TMEM lane quarter) does the gate math of the tiles
// of accumulator g while the MMA warp fills the other one.  With 18 - 72 K blocks per tile (bf16), the per-tile prologue
// (barrier setup, TMEM allocation, pipeline fill) and the un-overlapped epilogue were ~30 % of the one-tile-per-CTA kernel.
constexpr int kPCStages = 4;
constexpr int kPCSmem = 1024 + kPCStages * (kCABytes + kCBBytes) + 256;
constexpr int kPCThreads = 320;

// MC: clusters of two CTAs take two neighbouring pixel patches of the same (channel chunk, sample) in lockstep; each loads its
// own A tile and HALF of the weight tile, which TMA multicasts into both CTAs' shared memory: 256 instead of 384 128-byte rows
// requested per K block and SM (the L2 -> shared-memory fill, not the MMA rate, bounds this kernel: profiles/r02_convlstm_full.md).
// A stage is released to BOTH producers: the MMA commit arrives on the empty barrier of both CTAs (count 2).
template <bool BF16, bool MC>
__global__ void __launch_bounds__(kPCThreads, 1)
k_convlstm_tc_p(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias, const float* __restrict__ c_prev,
                float* __restrict__ h_out, __nv_bfloat16* __restrict__ h_bf, float* __restrict__ c_out, int H, int W, int C,
                int has_h, int tiles_w, int tiles_px, int tiles) {
    constexpr int kKE = BF16 ? 64 : kBlockK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kPCStages * kCABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kPCStages * kCBBytes);
    uint64_t* empty = full + kPCStages;
    uint64_t* acc_full = empty + kPCStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = C / kKE, nchunks = C / 64;
    const int kblocks = (has_h ? 2 : 1) * 9 * chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kPCStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MC ? 2 : 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kCN);
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();      // peers' barriers exist before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    // MC: the pair walks PAIR tiles (two neighbouring pixel patches); rank r takes patch 2 * pair_px + r (a patch index past
    // tiles_px is a dummy: TMA zero-fills it and the epilogue's bounds test drops it)
    const uint32_t crank = MC ? cluster_ctarank() : 0u;
    const int tile0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tstep = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int px_units = MC ? (tiles_px + 1) / 2 : tiles_px;

    if (warp == 0) {                                      // ===== TMA producer (whole warp converged, one lane issues) =====
        uint32_t s = 0, ph = 1;
        for (int tile = tile0; tile < tiles; tile += tstep) {
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = MC ? 2 * pu + (int)crank : pu;
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int h0 = th * kTH, w0 = tw * kTW;
            int src = 0, dy = -1, dx = -1, chunk = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], kCABytes + kCBBytes);
                    tma_load_4d(sA + s * kCABytes, src ? &tmH : &tmX, &full[s], chunk * kKE, w0 + dx, h0 + dy, b);
                    if (MC)
                        tma_load_2d_mc(sB + s * kCBBytes + crank * (kCBBytes / 2), &tmW, &full[s], kb * kKE,
                                       nchunk * kCN + (int)crank * (kCN / 2), (uint16_t)3);
                    else
                        tma_load_2d(sB + s * kCBBytes, &tmW, &full[s], kb * kKE, nchunk * kCN);
                }
                __syncwarp();
                if (++chunk == chunks) {
                    chunk = 0;
                    if (++dx == 2) { dx = -1; if (++dy == 2) { dy = -1; ++src; } }
                }
                if (++s == kPCStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {                               // ===== MMA issuer (whole warp converged, one lane issues) =====
        constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(128, kCN) : umma_idesc_tf32(128, kCN);
        uint32_t s = 0, ph = 0, lt = 0;
        for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
            const uint32_t buf = lt & 1;
            mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_acc + buf * kCN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_k128(smem_u32(sA + s * kCABytes));
                    const uint64_t db = umma_desc_k128(smem_u32(sB + s * kCBBytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        if (BF16) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma_tf32(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    if (MC) umma_commit_mc(&empty[s], (uint16_t)3);   // frees the stage in BOTH CTAs' producers' eyes
                    else umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == kPCStages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {                                              // ===== epilogue: group g = warps 2 + 4 g .. 5 + 4 g =====
        const int q = warp & 3;
        const uint32_t g = (uint32_t)(warp - 2) >> 2;
        uint32_t lt = 0;
        for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
            if ((lt & 1) != g) continue;
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = MC ? 2 * pu + (int)crank : pu;
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int r = q * 32 + lane;                  // accumulator row = pixel of the patch (h-major, w-minor)
            const int y = th * kTH + r / kTW, x = tw * kTW + r % kTW;
            const bool valid = y < H && x < W;
            const int64_t pix = (((int64_t)b * H + y) * W + x) * C + nchunk * 64;
            const float* bn = bias + nchunk * kCN;
            mbar_wait(&acc_full[g], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_acc + g * kCN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int sub = 0; sub < 4; ++sub) {           // 16 hidden channels at a time
                float gi[16], gf[16], go[16], gg[16];
                tmem_ld16_nowait(trow + 0 * 64 + sub * 16, gi);   // submodules.py:203 chunk order: in, remember, out, cell
                tmem_ld16_nowait(trow + 1 * 64 + sub * 16, gf);
                tmem_ld16_nowait(trow + 2 * 64 + sub * 16, go);
                tmem_ld16_nowait(trow + 3 * 64 + sub * 16, gg);
                tmem_ld_wait();
                if (sub == 3) {                           // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[g]);
                }
                if (valid) {
                    const int64_t e = pix + sub * 16;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c_prev) pc = *reinterpret_cast<const float4*>(c_prev + e + j);
                        const float4 bi = __ldg(reinterpret_cast<const float4*>(bn + 0 * 64 + sub * 16 + j));
                        const float4 bf = __ldg(reinterpret_cast<const float4*>(bn + 1 * 64 + sub * 16 + j));
                        const float4 bo = __ldg(reinterpret_cast<const float4*>(bn + 2 * 64 + sub * 16 + j));
                        const float4 bg = __ldg(reinterpret_cast<const float4*>(bn + 3 * 64 + sub * 16 + j));
                        float4 c, h;
                        c.x = act_sigm<BF16>(gf[j] + bf.x) * pc.x + act_sigm<BF16>(gi[j] + bi.x) * act_tanh<BF16>(gg[j] + bg.x);
                        c.y = act_sigm<BF16>(gf[j + 1] + bf.y) * pc.y + act_sigm<BF16>(gi[j + 1] + bi.y) * act_tanh<BF16>(gg[j + 1] + bg.y);
                        c.z = act_sigm<BF16>(gf[j + 2] + bf.z) * pc.z + act_sigm<BF16>(gi[j + 2] + bi.z) * act_tanh<BF16>(gg[j + 2] + bg.z);
                        c.w = act_sigm<BF16>(gf[j + 3] + bf.w) * pc.w + act_sigm<BF16>(gi[j + 3] + bi.w) * act_tanh<BF16>(gg[j + 3] + bg.w);
                        h.x = act_sigm<BF16>(go[j] + bo.x) * act_tanh<BF16>(c.x);
                        h.y = act_sigm<BF16>(go[j + 1] + bo.y) * act_tanh<BF16>(c.y);
                        h.z = act_sigm<BF16>(go[j + 2] + bo.z) * act_tanh<BF16>(c.z);
                        h.w = act_sigm<BF16>(go[j + 3] + bo.w) * act_tanh<BF16>(c.w);
                        *reinterpret_cast<float4*>(c_out + e + j) = c;
                        if (h_out) *reinterpret_cast<float4*>(h_out + e + j) = h;
                        if (BF16) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(h_bf + e + j) = pk;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    if (MC) cluster_sync_all(); else __syncthreads();      // no CTA leaves while its peer can still write into it
    if (warp == 1) tmem_dealloc(tmem_acc, 2 * kCN);
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2, OESS_CONVLSTM_2SM=1): the two CTAs of a cluster compute ONE 256-pixel x 256-column tile per
// step -- two neighbouring pixel patches x all four gates of a channel chunk.  Each CTA stages its own A box (16 KB) and HALF of
// the weight tile (16 KB) per K block, so the 192 KB ring holds SIX K blocks in flight instead of four (the kernel is bound by
// operand-fill latency x ring depth: profiles/r02_convlstm_full.md); the leader issues `tcgen05.mma.cta_group::2` with M = 256
// and commits to both CTAs' barriers; every CTA's epilogue works on its own 128 TMEM lanes as before.
constexpr int kP2Stages = 6;
constexpr int kP2BBytes = kCBBytes / 2;
constexpr int kP2Smem = 1024 + kP2Stages * (kCABytes + kP2BBytes) + 256;
constexpr int kP2Threads = 64 + 16 * 32;     // TMA warp, MMA warp, SIXTEEN epilogue warps (two per TMEM lane quarter and accumulator)

template <bool BF16>
__global__ void __launch_bounds__(kP2Threads, 1)
k_convlstm_tc_p2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                 const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias, const float* __restrict__ c_prev,
                 float* __restrict__ h_out, __nv_bfloat16* __restrict__ h_bf, float* __restrict__ c_out, int H, int W, int C,
                 int has_h, int tiles_w, int tiles_px, int tiles) {
    constexpr int kKE = BF16 ? 64 : kBlockK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kP2Stages * kCABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kP2Stages * kP2BBytes);
    uint64_t* empty = full + kP2Stages;
    uint64_t* acc_full = empty + kP2Stages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = C / kKE, nchunks = C / 64;
    const int kblocks = (has_h ? 2 : 1) * 9 * chunks;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kP2Stages; ++s) {
            mbar_init(&full[s], 1);                       // the leader's own arrive.expect_tx (both CTAs' bytes)
            mbar_init(&empty[s], 1);                      // the leader's multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);                   // the leader's multicast commit
            mbar_init(&acc_empty[b], 16);                 // eight epilogue warps per accumulator in each CTA (used in the leader only)
        }
        mbar_fence_init();
    }
    cluster_sync_all();                                   // barriers of both CTAs exist before TMEM allocation / any remote arrive
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * kCN);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const int tile0 = (int)(blockIdx.x >> 1), tstep = (int)(gridDim.x >> 1);
    const int px_units = (tiles_px + 1) / 2;

    if (warp == 0) {                                      // ===== TMA producer (both CTAs) =====
        uint32_t s = 0, ph = 1;
        for (int tile = tile0; tile < tiles; tile += tstep) {
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = 2 * pu + (int)crank;           // past tiles_px: a dummy patch (TMA zero fill, epilogue drops it)
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int h0 = th * kTH, w0 = tw * kTW;
            int src = 0, dy = -1, dx = -1, chunk = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[s], 2 * (kCABytes + kP2BBytes));
                    tma_load_4d_2sm(sA + s * kCABytes, src ? &tmH : &tmX, &full[s], chunk * kKE, w0 + dx, h0 + dy, b);
                    tma_load_2d_2sm(sB + s * kP2BBytes, &tmW, &full[s], kb * kKE, nchunk * kCN + (int)crank * (kCN / 2));
                }
                __syncwarp();
                if (++chunk == chunks) {
                    chunk = 0;
                    if (++dx == 2) { dx = -1; if (++dy == 2) { dy = -1; ++src; } }
                }
                if (++s == kP2Stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {                               // ===== MMA issuer (leader CTA only) =====
        if (leader) {
            constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(256, kCN) : umma_idesc_tf32(256, kCN);
            uint32_t s = 0, ph = 0, lt = 0;
            for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
                const uint32_t buf = lt & 1;
                mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_acc + buf * kCN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = umma_desc_k128(smem_u32(sA + s * kCABytes));
                        const uint64_t db = umma_desc_k128(smem_u32(sB + s * kP2BBytes));
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            if (BF16) umma_bf16_2sm(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            else umma_tf32_2sm(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit_2sm(&empty[s], (uint16_t)3);
                    }
                    __syncwarp();
                    if (++s == kP2Stages) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(&acc_full[buf], (uint16_t)3);
                __syncwarp();
            }
        }
    } else {                                              // ===== epilogue: 16 warps (both CTAs) =====
        // warp e = warp - 2: TMEM lane quarter q = warp & 3, channel half hs = (e >> 2) & 1 (hidden channels 32 hs .. 32 hs + 31 of the
        // chunk), accumulator g = e >> 3.  Four epilogue warps per SM sub-partition instead of two: at the C = 64 level the MMA
        // warp still waited 36 % of its time for a drained accumulator (profiles/r02_convlstm2sm_full.md).
        const int e = warp - 2;
        const int q = warp & 3, hs = (e >> 2) & 1;       // a warp may only touch the TMEM lanes of quarter warp % 4
        const uint32_t g = (uint32_t)e >> 3;
        uint32_t lt = 0;
        for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
            if ((lt & 1) != g) continue;
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = 2 * pu + (int)crank;
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int r = q * 32 + lane;
            const int y = th * kTH + r / kTW, x = tw * kTW + r % kTW;
            const bool valid = y < H && x < W;
            const int64_t pix = (((int64_t)b * H + y) * W + x) * C + nchunk * 64 + hs * 32;
            const float* bn = bias + nchunk * kCN + hs * 32;
            mbar_wait(&acc_full[g], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_acc + g * kCN + ((uint32_t)(q * 32) << 16) + (uint32_t)(hs * 32);
#pragma unroll 1
            for (int sub = 0; sub < 4; ++sub) {           // 8 hidden channels at a time
                float gi[8], gf[8], go[8], gg[8];
                tmem_ld8_nowait(trow + 0 * 64 + sub * 8, gi);     // submodules.py:203 chunk order: in, remember, out, cell
                tmem_ld8_nowait(trow + 1 * 64 + sub * 8, gf);
                tmem_ld8_nowait(trow + 2 * 64 + sub * 8, go);
                tmem_ld8_nowait(trow + 3 * 64 + sub * 8, gg);
                tmem_ld_wait();
                if (sub == 3) {                           // this warp's part of the accumulator is read: tell the LEADER's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&acc_empty[g], 0);
                }
                if (valid) {
                    const int64_t e0 = pix + sub * 8;
#pragma unroll
                    for (int j = 0; j < 8; j += 4) {
                        float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c_prev) pc = *reinterpret_cast<const float4*>(c_prev + e0 + j);
                        const float4 bi = __ldg(reinterpret_cast<const float4*>(bn + 0 * 64 + sub * 8 + j));
                        const float4 bf = __ldg(reinterpret_cast<const float4*>(bn + 1 * 64 + sub * 8 + j));
                        const float4 bo = __ldg(reinterpret_cast<const float4*>(bn + 2 * 64 + sub * 8 + j));
                        const float4 bg = __ldg(reinterpret_cast<const float4*>(bn + 3 * 64 + sub * 8 + j));
                        float4 c, h;
                        c.x = act_sigm<BF16>(gf[j] + bf.x) * pc.x + act_sigm<BF16>(gi[j] + bi.x) * act_tanh<BF16>(gg[j] + bg.x);
                        c.y = act_sigm<BF16>(gf[j + 1] + bf.y) * pc.y + act_sigm<BF16>(gi[j + 1] + bi.y) * act_tanh<BF16>(gg[j + 1] + bg.y);
                        c.z = act_sigm<BF16>(gf[j + 2] + bf.z) * pc.z + act_sigm<BF16>(gi[j + 2] + bi.z) * act_tanh<BF16>(gg[j + 2] + bg.z);
                        c.w = act_sigm<BF16>(gf[j + 3] + bf.w) * pc.w + act_sigm<BF16>(gi[j + 3] + bi.w) * act_tanh<BF16>(gg[j + 3] + bg.w);
                        h.x = act_sigm<BF16>(go[j] + bo.x) * act_tanh<BF16>(c.x);
                        h.y = act_sigm<BF16>(go[j + 1] + bo.y) * act_tanh<BF16>(c.y);
                        h.z = act_sigm<BF16>(go[j + 2] + bo.z) * act_tanh<BF16>(c.z);
                        h.w = act_sigm<BF16>(go[j + 3] + bo.w) * act_tanh<BF16>(c.w);
                        *reinterpret_cast<float4*>(c_out + e0 + j) = c;
                        if (h_out) *reinterpret_cast<float4*>(h_out + e0 + j) = h;
                        if (BF16) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(h_bf + e0 + j) = pk;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its peer still works
    if (warp == 1) tmem_dealloc_2sm(tmem_acc, 2 * kCN);
}

// ---------------------------------------------------------------------------------------------------------------------
// Haloed CTA-pair variant (default where its patches tile the image as well; OESS_CONVLSTM_HALO=0 / 1): the nine taps of a (source, channel chunk) read ONE 18-row x 10-pixel halo
// box (23 KB) instead of nine 16 KB boxes.  The pixel patch of a CTA is 8 wide x 16 high, so every 8-row core-matrix group of
// the A operand is one patch row and a tap's shifted window is the halo tile read from row ky * 10 + kx on with 8-row groups
// 10 rows (1 280 B = the descriptor's stride byte offset) apart; the 128-byte swizzle is a function of the shared-memory
// address, so a start address that is not a multiple of 8 rows reads what TMA wrote (tools/probe_umma_window.cu checks exactly
// this on the device).  Fill per tile and CTA at C = 64: 2 x 23 KB of A + 18 x 16 KB of B = 334 KB instead of 576 KB (the pair
// kernel above sits on the L2 -> shared-memory fill bound).  A halos and B boxes run in separate rings.
constexpr int kHW = 8, kHH = 16;                          // pixel patch of one CTA = the 128 rows of M
constexpr int kHaloW = kHW + 2, kHaloH = kHH + 2;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;          // 23 040
constexpr int kHaloSlot = 23 * 1024;
constexpr int kH2AStages = 3, kH2BStages = 9;
constexpr int kH2Smem = 1024 + kH2AStages * kHaloSlot + kH2BStages * kP2BBytes + 256;

template <bool BF16>
__global__ void __launch_bounds__(kP2Threads, 1)
k_convlstm_tc_h2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                 const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias, const float* __restrict__ c_prev,
                 float* __restrict__ h_out, __nv_bfloat16* __restrict__ h_bf, float* __restrict__ c_out, int H, int W, int C,
                 int has_h, int tiles_w, int tiles_px, int tiles) {
    constexpr int kKE = BF16 ? 64 : kBlockK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kH2AStages * kHaloSlot;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + kH2BStages * kP2BBytes);
    uint64_t* emptyA = fullA + kH2AStages;
    uint64_t* fullB = emptyA + kH2AStages;
    uint64_t* emptyB = fullB + kH2BStages;
    uint64_t* acc_full = emptyB + kH2BStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = C / kKE, nchunks = C / 64;
    const int groups = (has_h ? 2 : 1) * chunks;          // (source, channel chunk) pairs per tile
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kH2AStages; ++s) {
            mbar_init(&fullA[s], 1);
            mbar_init(&emptyA[s], 1);
        }
        for (int s = 0; s < kH2BStages; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 16);
        }
        mbar_fence_init();
    }
    cluster_sync_all();
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * kCN);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const int tile0 = (int)(blockIdx.x >> 1), tstep = (int)(gridDim.x >> 1);
    const int px_units = (tiles_px + 1) / 2;

    if (warp == 0) {                                      // ===== TMA producer (both CTAs) =====
        uint32_t sa = 0, pha = 1, sb = 0, phb = 1;
        for (int tile = tile0; tile < tiles; tile += tstep) {
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = 2 * pu + (int)crank;           // past tiles_px: a dummy patch (TMA zero fill, epilogue drops it)
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int h0 = th * kHH, w0 = tw * kHW;
            for (int g = 0; g < groups; ++g) {
                const int src = g / chunks, chunk = g - src * chunks;
                mbar_wait(&emptyA[sa], pha);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&fullA[sa], 2 * kHaloBytes);
                    tma_load_4d_2sm(sA + sa * kHaloSlot, src ? &tmH : &tmX, &fullA[sa], chunk * kKE, w0 - 1, h0 - 1, b);
                }
                __syncwarp();
                if (++sa == kH2AStages) { sa = 0; pha ^= 1; }
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(&emptyB[sb], phb);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(&fullB[sb], 2 * kP2BBytes);
                        tma_load_2d_2sm(sB + sb * kP2BBytes, &tmW, &fullB[sb], ((src * 9 + tap) * chunks + chunk) * kKE,
                                        nchunk * kCN + (int)crank * (kCN / 2));
                    }
                    __syncwarp();
                    if (++sb == kH2BStages) { sb = 0; phb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {                               // ===== MMA issuer (leader CTA only) =====
        if (leader) {
            constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(256, kCN) : umma_idesc_tf32(256, kCN);
            uint32_t sa = 0, pha = 0, sb = 0, phb = 0, lt = 0;
            for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
                const uint32_t buf = lt & 1;
                mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_acc + buf * kCN;
                for (int g = 0; g < groups; ++g) {
                    mbar_wait(&fullA[sa], pha);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + sa * kHaloSlot);
                    int ky = 0, kx = 0;
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&fullB[sb], phb);
                        tc_fence_after();
                        if (elect_one()) {
                            // window of tap (ky, kx): halo rows (r + ky) * 10 + kx + c, r = patch row = 8-row group, c = 0..7
                            const uint32_t a_addr = a0 + (uint32_t)(ky * kHaloW + kx) * 128u;
                            const uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFFu) | (1ull << 16) |
                                                ((uint64_t)((kHaloW * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
                            const uint64_t db = umma_desc_k128(smem_u32(sB + sb * kP2BBytes));
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                if (BF16) umma_bf16_2sm(d, da + 2 * k, db + 2 * k, idesc, (g | tap | k) != 0);
                                else umma_tf32_2sm(d, da + 2 * k, db + 2 * k, idesc, (g | tap | k) != 0);
                            }
                            umma_commit_2sm(&emptyB[sb], (uint16_t)3);
                            if (tap == 8) umma_commit_2sm(&emptyA[sa], (uint16_t)3);
                        }
                        __syncwarp();
                        if (++kx == 3) { kx = 0; ++ky; }
                        if (++sb == kH2BStages) { sb = 0; phb ^= 1; }
                    }
                    if (++sa == kH2AStages) { sa = 0; pha ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(&acc_full[buf], (uint16_t)3);
                __syncwarp();
            }
        }
    } else {                                              // ===== epilogue: 16 warps (both CTAs), as in k_convlstm_tc_p2 =====
        const int e = warp - 2;
        const int q = warp & 3, hs = (e >> 2) & 1;
        const uint32_t g = (uint32_t)e >> 3;
        uint32_t lt = 0;
        for (int tile = tile0; tile < tiles; tile += tstep, ++lt) {
            if ((lt & 1) != g) continue;
            const int pu = tile % px_units, rest = tile / px_units;
            const int px = 2 * pu + (int)crank;
            const int nchunk = rest % nchunks, b = rest / nchunks;
            const int th = px / tiles_w, tw = px - th * tiles_w;
            const int r = q * 32 + lane;
            const int y = th * kHH + r / kHW, x = tw * kHW + r % kHW;
            const bool valid = y < H && x < W;
            const int64_t pix = (((int64_t)b * H + y) * W + x) * C + nchunk * 64 + hs * 32;
            const float* bn = bias + nchunk * kCN + hs * 32;
            mbar_wait(&acc_full[g], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_acc + g * kCN + ((uint32_t)(q * 32) << 16) + (uint32_t)(hs * 32);
#pragma unroll 1
            for (int sub = 0; sub < 4; ++sub) {
                float gi[8], gf[8], go[8], gg[8];
                tmem_ld8_nowait(trow + 0 * 64 + sub * 8, gi);
                tmem_ld8_nowait(trow + 1 * 64 + sub * 8, gf);
                tmem_ld8_nowait(trow + 2 * 64 + sub * 8, go);
                tmem_ld8_nowait(trow + 3 * 64 + sub * 8, gg);
                tmem_ld_wait();
                if (sub == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&acc_empty[g], 0);
                }
                if (valid) {
                    const int64_t e0 = pix + sub * 8;
#pragma unroll
                    for (int j = 0; j < 8; j += 4) {
                        float4 pc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c_prev) pc = *reinterpret_cast<const float4*>(c_prev + e0 + j);
                        const float4 bi = __ldg(reinterpret_cast<const float4*>(bn + 0 * 64 + sub * 8 + j));
                        const float4 bf = __ldg(reinterpret_cast<const float4*>(bn + 1 * 64 + sub * 8 + j));
                        const float4 bo = __ldg(reinterpret_cast<const float4*>(bn + 2 * 64 + sub * 8 + j));
                        const float4 bg = __ldg(reinterpret_cast<const float4*>(bn + 3 * 64 + sub * 8 + j));
                        float4 c, h;
                        c.x = act_sigm<BF16>(gf[j] + bf.x) * pc.x + act_sigm<BF16>(gi[j] + bi.x) * act_tanh<BF16>(gg[j] + bg.x);
                        c.y = act_sigm<BF16>(gf[j + 1] + bf.y) * pc.y + act_sigm<BF16>(gi[j + 1] + bi.y) * act_tanh<BF16>(gg[j + 1] + bg.y);
                        c.z = act_sigm<BF16>(gf[j + 2] + bf.z) * pc.z + act_sigm<BF16>(gi[j + 2] + bi.z) * act_tanh<BF16>(gg[j + 2] + bg.z);
                        c.w = act_sigm<BF16>(gf[j + 3] + bf.w) * pc.w + act_sigm<BF16>(gi[j + 3] + bi.w) * act_tanh<BF16>(gg[j + 3] + bg.w);
                        h.x = act_sigm<BF16>(go[j] + bo.x) * act_tanh<BF16>(c.x);
                        h.y = act_sigm<BF16>(go[j + 1] + bo.y) * act_tanh<BF16>(c.y);
                        h.z = act_sigm<BF16>(go[j + 2] + bo.z) * act_tanh<BF16>(c.z);
                        h.w = act_sigm<BF16>(go[j + 3] + bo.w) * act_tanh<BF16>(c.w);
                        *reinterpret_cast<float4*>(c_out + e0 + j) = c;
                        if (h_out) *reinterpret_cast<float4*>(h_out + e0 + j) = h;
                        if (BF16) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(h_bf + e0 + j) = pk;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_acc, 2 * kCN);
}

}  // namespace tc
}  // namespace oess

using namespace oess;

// x, h_prev, c_prev, h_out, c_out: [B, H, W, C] channels-last float32 (h_prev / c_prev NULL = zero state).
// w_packed: [4C, 2 * 9 * C] with row n' = chunk * 256 + gate * 64 + c  (hidden channel chunk * 64 + c) and column
// (source, tap = ky * 3 + kx, channel); bias_packed: [4C] in the same row order.  C % 64 == 0.
template <bool BF16>
static int convlstm_impl(const void* x, const void* h_prev, const float* c_prev, const void* w_packed, const float* bias_packed,
                         float* h_out, __nv_bfloat16* h_bf, float* c_out, int B, int H, int W, int C, oess_stream_t stream) {
    constexpr int ES = BF16 ? 2 : 4;                      // operand element size
    constexpr uint32_t KE = BF16 ? 64 : tc::kBlockK;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 64) != 0) return OESS_E_ARG;
    if (!x || !w_packed || !bias_packed || !c_out || (BF16 ? !h_bf : !h_out)) return OESS_E_ARG;
    if (((uintptr_t)x | (uintptr_t)h_prev | (uintptr_t)c_prev | (uintptr_t)w_packed | (uintptr_t)bias_packed |
         (uintptr_t)h_out | (uintptr_t)h_bf | (uintptr_t)c_out) & 15)
        return OESS_E_ARG;
    if (B > 65535 || C / 64 > 65535) return OESS_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    // w_packed always holds both sources' columns ([4C, 2 * 9 * C]); with a zero hidden state the kernel reads only the
    // first 9 * C columns of every row (source 0), so the tensor map keeps the full row stride.
    const uint64_t Kfull = (uint64_t)2 * 9 * C;
    CUtensorMap tmX, tmH, tmW;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * ES, (uint64_t)W * C * ES, (uint64_t)H * W * C * ES};
    const uint32_t box[4] = {KE, tc::kTW, tc::kTH, 1};
    auto mk = BF16 ? tc::make_tmap_bf16 : tc::make_tmap_f32;
    int rc = mk(&tmX, x, 4, dims, strides, box);
    if (rc) return rc;
    rc = mk(&tmH, h_prev ? h_prev : x, 4, dims, strides, box);
    if (rc) return rc;
    const uint64_t dW[2] = {Kfull, (uint64_t)4 * C}, sW[1] = {Kfull * ES};
    // haloed pair kernel: default where its 8 x 16 patches tile the image with no more patches than the 16 x 8 ones of the other
    // kernels (55 x 80: 40 against 35 -- slower there); OESS_CONVLSTM_HALO=0 never, =1 always
    static const int halo_env = [] { const char* e = std::getenv("OESS_CONVLSTM_HALO"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
    const int64_t halo_px = (int64_t)((W + tc::kHW - 1) / tc::kHW) * ((H + tc::kHH - 1) / tc::kHH);
    const int64_t tile_px = (int64_t)((W + tc::kTW - 1) / tc::kTW) * ((H + tc::kTH - 1) / tc::kTH);
    if (halo_px >= 2 && (halo_env == 1 || (halo_env < 0 && halo_px <= tile_px))) {
        CUtensorMap tmXh, tmHh, tmWh;
        const uint32_t hbox[4] = {KE, (uint32_t)tc::kHaloW, (uint32_t)tc::kHaloH, 1};
        rc = mk(&tmXh, x, 4, dims, strides, hbox);
        if (rc) return rc;
        rc = mk(&tmHh, h_prev ? h_prev : x, 4, dims, strides, hbox);
        if (rc) return rc;
        const uint32_t bWh[2] = {KE, (uint32_t)(tc::kCN / 2)};
        rc = mk(&tmWh, w_packed, 2, dW, sW, bWh);
        if (rc) return rc;
        const int htw = (W + tc::kHW - 1) / tc::kHW, hth = (H + tc::kHH - 1) / tc::kHH;
        const int hpx = htw * hth;
        const int64_t htiles = (int64_t)((hpx + 1) / 2) * (C / 64) * B;
        if (htiles < (1ll << 31)) {
            auto kern = tc::k_convlstm_tc_h2<BF16>;
            OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kH2Smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(2 * (htiles < kNumSMs / 2 ? htiles : kNumSMs / 2)));
            cfg.blockDim = dim3(tc::kP2Threads);
            cfg.dynamicSmemBytes = tc::kH2Smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            OESS_KERNEL(BF16 ? "tc_convlstm_step_bf16" : "tc_convlstm_step", st,
                        cudaLaunchKernelEx(&cfg, kern, tmXh, tmHh, tmWh, bias_packed, c_prev, h_out, h_bf, c_out, H, W, C,
                                           h_prev ? 1 : 0, htw, hpx, (int)htiles));
            return 0;
        }
    }
    const int tiles_w = (W + tc::kTW - 1) / tc::kTW, tiles_h = (H + tc::kTH - 1) / tc::kTH;
    static const bool tile_env = [] { const char* e = std::getenv("OESS_CONVLSTM"); return e && e[0] == 't'; }();
    static const bool mc_env = [] { const char* e = std::getenv("OESS_CONVLSTM_MC"); return !(e && e[0] == '0'); }();   // default on
    const int tiles_px = tiles_w * tiles_h;
    static const bool sm2_env = [] { const char* e = std::getenv("OESS_CONVLSTM_2SM"); return !(e && e[0] == '0'); }();   // default on
    const bool sm2 = sm2_env && !tile_env && tiles_px >= 2;
    const bool mc = (mc_env || sm2) && !tile_env && tiles_px >= 2;
    const uint32_t bW[2] = {KE, (uint32_t)(mc ? tc::kCN / 2 : tc::kCN)};
    rc = mk(&tmW, w_packed, 2, dW, sW, bW);
    if (rc) return rc;
    const int64_t tiles = (int64_t)(mc ? (tiles_px + 1) / 2 : tiles_px) * (C / 64) * B;
    if (!tile_env && tiles < (1ll << 31)) {
        auto kern = sm2 ? tc::k_convlstm_tc_p2<BF16> : (mc ? tc::k_convlstm_tc_p<BF16, true> : tc::k_convlstm_tc_p<BF16, false>);
        const int smem_bytes = sm2 ? tc::kP2Smem : tc::kPCSmem;
        OESS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
        if (mc) grid = (unsigned)(2 * (tiles < kNumSMs / 2 ? tiles : kNumSMs / 2));      // whole clusters
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(sm2 ? tc::kP2Threads : tc::kPCThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = mc ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        OESS_KERNEL(BF16 ? "tc_convlstm_step_bf16" : "tc_convlstm_step", st,
                    cudaLaunchKernelEx(&cfg, kern, tmX, tmH, tmW, bias_packed, c_prev, h_out, h_bf, c_out, H, W, C,
                                       h_prev ? 1 : 0, tiles_w, tiles_px, (int)tiles));
        return 0;
    }
    OESS_CUDA(cudaFuncSetAttribute(tc::k_convlstm_tc<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kCSmem));
    const dim3 grid((unsigned)(tiles_w * tiles_h), (unsigned)(C / 64), (unsigned)B);
    OESS_KERNEL(BF16 ? "tc_convlstm_step_bf16" : "tc_convlstm_step", st, tc::k_convlstm_tc<BF16><<<grid, 192, tc::kCSmem, st>>>(
        tmX, tmH, tmW, bias_packed, c_prev, h_out, h_bf, c_out, H, W, C, h_prev ? 1 : 0, tiles_w));
    return 0;
}

OESS_API int oess_convlstm_step_nhwc(const float* x, const float* h_prev, const float* c_prev, const float* w_packed,
                                     const float* bias_packed, float* h_out, float* c_out, int B, int H, int W, int C,
                                     oess_stream_t stream) {
    return convlstm_impl<false>(x, h_prev, c_prev, w_packed, bias_packed, h_out, nullptr, c_out, B, H, W, C, stream);
}

// bf16-operand variant for the FROZEN E2VID encoder: x, h_prev, w_packed are bfloat16 (same layouts), the cell state and the
// bias stay fp32, accumulation is fp32.  h_bf16_out (required) is the hidden state as the next step's operand; h_out
// (fp32, may be NULL) is the same state for fp32 consumers (the next encoder level, the latent dictionary).
OESS_API int oess_convlstm_step_nhwc_bf16(const void* x, const void* h_prev, const float* c_prev, const void* w_packed,
                                          const float* bias_packed, float* h_out, void* h_bf16_out, float* c_out, int B, int H,
                                          int W, int C, oess_stream_t stream) {
    return convlstm_impl<true>(x, h_prev, c_prev, w_packed, bias_packed, h_out, (__nv_bfloat16*)h_bf16_out, c_out, B, H, W, C,
                               stream);
}
