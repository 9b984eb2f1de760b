// tri_band.cuh -- banded ORDERED path of the trilinear voxeliser (steps 2 and 3; step 1 is the row radix pass).
//
// Input: per frame, 16-byte records (x, y, t_norm, value) stably sorted by source-cell row py = y0 + 1, with
// rowoff[f][py] = first record of row py (the radix pass's scanned bin totals).  Events that can never touch
// the grid (y out of reach, t_norm NaN/out of int range) were keyed into the sentinel row H + 1 and are
// never read here.
#pragma once
#include "radix.cuh"
#include "tri_common.cuh"

namespace oess {
namespace tri {

// ---------------------------------------------------------------------------------------------
// Step 2: stable counting sort of one (frame, row) segment by source-cell column.
// One warp per row; chunks of 32 events are ranked in order (match_any + popc), so ties keep event
// order.  Records whose column cannot reach the grid get the canonical x = -8 (they sort last in the row
// and never pass the 0 <= xl < W test), so the splat kernel can use plain float->int conversions.
// rowflag[f][py] = 1 if, inside some cell of the row, the time bins t0 are not non-decreasing in event
// order (only possible for input that is not time-sorted): the splat kernel then takes its robust path.
// ---------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;
constexpr float kCanonicalBadX = -8.0f;

__global__ void __launch_bounds__(kRowWarps * 32)
k_rowsort(const float4* __restrict__ src, float4* __restrict__ dst, const int64_t* __restrict__ frame_offsets,
          const uint32_t* __restrict__ rowoff, uint32_t* __restrict__ rowflag, int H, int W,
          uint32_t* __restrict__ coloff, int NS, int WC) {
    extern __shared__ uint32_t s_cnt_all[];           // [kRowWarps][W + 2]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.y * kRowWarps + w;        // py in [0, H]
    const int f = blockIdx.x;                          // frame-major launch order: heavy and light frames interleave
    if (row > H) return;
    const uint32_t* ro = rowoff + (int64_t)f * radix::kBins;
    const uint32_t s = ro[row], e = ro[row + 1];
    // strip formulation (tri_strip.cuh): coloff[f][row][2 k + j] = first record with column >= min(k WC, W) + j
    const int KT = 2 * (NS + 1);
    uint32_t* T = coloff ? coloff + ((int64_t)f * (H + 1) + row) * KT : nullptr;
    if (s == e) {
        if (lane == 0) rowflag[(int64_t)f * radix::kBins + row] = 0;
        if (T) for (int k = lane; k < KT; k += 32) T[k] = s;
        return;
    }
    const int nb = W + 2;
    uint32_t* cnt = s_cnt_all + w * nb;
    const float4* in = src + frame_offsets[f];
    float4* out = dst + frame_offsets[f];
    for (int b = lane; b < nb; b += 32) cnt[b] = 0;
    __syncwarp();
    for (uint32_t i = s + lane; i < e; i += 32) atomicAdd(&cnt[cell_px(in[i].x, W)], 1u);
    __syncwarp();
    // exclusive scan of the W + 2 bins: each lane owns an odd-length (bank-conflict-free) segment of bins,
    // sums it serially, one warp scan of the 32 partials, then rewrites its segment
    {
        const int seg = ((nb + 31) / 32) | 1;
        const int b_lo = min(lane * seg, nb), b_hi = min(b_lo + seg, nb);
        uint32_t sum = 0;
        for (int b = b_lo; b < b_hi; ++b) sum += cnt[b];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        uint32_t run = s + incl - sum;
        for (int b = b_lo; b < b_hi; ++b) {
            const uint32_t v = cnt[b];
            cnt[b] = run;
            run += v;
        }
    }
    __syncwarp();
    if (T) for (int k = lane; k < KT; k += 32) T[k] = cnt[min((k >> 1) * WC, W) + (k & 1)];
    __syncwarp();
    const unsigned lt = lanemask_lt();
    for (uint32_t i0 = s; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool act = i < e;
        float4 it = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t px = 0xffffffffu;
        if (act) {
            it = in[i];
            px = cell_px(it.x, W);
            if (px > (uint32_t)W) it.x = kCanonicalBadX;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, px);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader && act) { base = cnt[px]; cnt[px] = base + __popc(peers); }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (act) out[base + __popc(peers & lt)] = it;
        __syncwarp();
    }
    // time-bin monotonicity inside every cell of the sorted row
    bool bad = false;
    for (uint32_t i = s + 1 + lane; i < e; i += 32) {
        const float4 a = out[i - 1], b = out[i];
        bad |= (__float2int_rz(a.x) == __float2int_rz(b.x)) && (__float2int_rz(b.z) < __float2int_rz(a.z));
    }
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) rowflag[(int64_t)f * radix::kBins + row] = anybad ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------------
// Step 3: ordered splat of one band of TH output rows from cell-sorted records.
//
// The CTA keeps the band's C x TH x W accumulators in shared memory.  The band's records (source rows
// ty0-1 .. ty0+TH-1, one contiguous range) are cut into 8 per-warp chunks whose ends are moved to cell-run
// boundaries, so every cell run is replayed by exactly one warp.  The four (dx, dy) passes run in reference
// order (representations.py:33-34) separated by __syncthreads(); inside one pass different cells hit
// different pixel columns, so warps never conflict.  LANES = EVENTS (dense: pixels are 3x more numerous than
// events and mostly empty).  Lanes that hit the same accumulator (same cell, same time bin) add one per
// round, in event order, with a plain shared-memory read-add-write; the dt = 0 adds of a cell precede its
// dt = 1 adds (:35).
//   fast path  : time bins are non-decreasing inside every cell (always true for time-sorted events), so
//                same-accumulator lanes are adjacent and ranks come from one ballot; windows are cut at run
//                boundaries so that one window does dt = 0 then dt = 1 from a single load of the record.
//   robust path: any input; separate dt sweeps, ranks from match_any.
// The band is written out once, coalesced (no memset of the grid, no global atomics).
// ---------------------------------------------------------------------------------------------
constexpr int kBandThreads = 256;

__device__ __forceinline__ uint32_t run_key(const float4& r) {   // canonical records only
    return ((uint32_t)(__float2int_rz(r.y) + 1) << 16) | (uint32_t)(__float2int_rz(r.x) + 8);
}

// first index i in [s, e_hi] at which a cell run starts (i == e_lo, i == e_hi, or key(i) != key(i-1))
__device__ __forceinline__ uint32_t run_start_at_or_after(const float4* __restrict__ it, uint32_t s, uint32_t e_lo,
                                                          uint32_t e_hi, int lane) {
    if (s <= e_lo) return e_lo;
    if (s >= e_hi) return e_hi;
    uint32_t prev = run_key(it[s - 1]);
    for (uint32_t base = s; base < e_hi; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t k = (i < e_hi) ? run_key(it[i]) : 0xffffffffu;
        uint32_t p = __shfl_up_sync(0xffffffffu, k, 1);
        if (lane == 0) p = prev;
        const unsigned heads = __ballot_sync(0xffffffffu, k != p);
        if (heads) return base + (uint32_t)(__ffs(heads) - 1);
        prev = __shfl_sync(0xffffffffu, k, 31);
    }
    return e_hi;
}

struct BandCtx {
    float* acc;
    int dx, dy, ty0, rows, C, TH, W, lane;
};

// one record, one dt: target accumulator (or a unique per-lane sentinel) and corner weight
__device__ __forceinline__ uint32_t corner(const BandCtx& c, const float4& r, bool on, int dt, float* wgt) {
    const int xl = __float2int_rz(r.x) + c.dx;         // representations.py:27,33
    const int yl = __float2int_rz(r.y) + c.dy;         // :28,34
    const int tl = __float2int_rz(r.z) + dt;           // :29,35
    const int yrel = yl - c.ty0;
    *wgt = weight_t(weight_xy(r.x, r.y, r.w, xl, yl), r.z, tl);                                   // :37
    const bool ok = on && (unsigned)xl < (unsigned)c.W && (unsigned)yrel < (unsigned)c.rows &&
                    (unsigned)tl < (unsigned)c.C;                                                 // :36
    return ok ? (uint32_t)((tl * c.TH + yrel) * c.W + xl) : (0x80000000u | (uint32_t)c.lane);
}

// ordered accumulate of one warp-wide slice; same-address lanes are ADJACENT (fast path)
__device__ __forceinline__ void commit_adjacent(const BandCtx& c, uint32_t addr, float wgt) {
    const bool ok = !(addr & 0x80000000u);
    const uint32_t prev = __shfl_up_sync(0xffffffffu, addr, 1);
    const bool head = (c.lane == 0) || addr != prev;
    const unsigned dup = __ballot_sync(0xffffffffu, ok && !head);
    if (dup == 0) {
        if (ok) c.acc[addr] = __fadd_rn(c.acc[addr], wgt);                                      // :43
    } else {
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        const int rank = c.lane - (31 - __clz(hm & (0xffffffffu >> (31 - c.lane))));
        const int rounds = __reduce_max_sync(0xffffffffu, ok ? rank + 1 : 0);
        for (int k = 0; k < rounds; ++k) {
            if (ok && rank == k) c.acc[addr] = __fadd_rn(c.acc[addr], wgt);
            __syncwarp();
        }
    }
    __syncwarp();
}

// ordered accumulate of one warp-wide slice; same-address lanes anywhere in the warp (robust path)
__device__ __forceinline__ void commit_any(const BandCtx& c, uint32_t addr, float wgt) {
    const bool ok = !(addr & 0x80000000u);
    const unsigned peers = __match_any_sync(0xffffffffu, addr);
    const int rank = __popc(peers & lanemask_lt());
    const int rounds = __reduce_max_sync(0xffffffffu, ok ? rank + 1 : 0);
    for (int k = 0; k < rounds; ++k) {
        if (ok && rank == k) c.acc[addr] = __fadd_rn(c.acc[addr], wgt);
        __syncwarp();
    }
}

// all events of [lo, hi): dt = 0 sweep then dt = 1 sweep (any run length, any time order)
template <bool ADJ>
__device__ __forceinline__ void sweep_two_pass(const BandCtx& c, const float4* __restrict__ it, uint32_t lo, uint32_t hi) {
#pragma unroll 1
    for (int dt = 0; dt < 2; ++dt) {
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t i = base + c.lane;
            const bool on = i < hi;
            const float4 r = on ? it[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float wgt;
            const uint32_t addr = corner(c, r, on, dt, &wgt);
            if (ADJ) commit_adjacent(c, addr, wgt); else commit_any(c, addr, wgt);
        }
    }
}

__global__ void __launch_bounds__(kBandThreads, 4)
k_band_splat(const float4* __restrict__ items, const int64_t* __restrict__ frame_offsets,
             const uint32_t* __restrict__ rowoff, const uint32_t* __restrict__ rowflag, Geom g, int TH,
             int stage_cap, int npass, float* __restrict__ out) {
    extern __shared__ __align__(16) float s_acc[];     // [C][TH][W]
    __shared__ int s_robust;
    const int f = blockIdx.y;
    const int ty0 = blockIdx.x * TH;
    const int rows = min(TH, g.H - ty0);
    const int W = g.W, C = g.C, H = g.H;
    const int nacc = C * TH * W;
    const bool vec4 = (W & 3) == 0;
    if (threadIdx.x == 0) s_robust = 0;
    if (vec4) {
        float4* a4 = reinterpret_cast<float4*>(s_acc);
        for (int i = threadIdx.x; i < nacc / 4; i += kBandThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int i = threadIdx.x; i < nacc; i += kBandThreads) s_acc[i] = 0.0f;   // representations.py:22 zeros
    }
    __syncthreads();

    // source-cell rows py in [ty0, ty0 + rows] feed output rows [ty0, ty0 + rows)
    const int py_hi = min(ty0 + rows, H);
    const uint32_t* ro = rowoff + (int64_t)f * radix::kBins;
    const uint32_t e_lo = ro[ty0], e_hi = ro[py_hi + 1];
    if (threadIdx.x <= py_hi - ty0 && rowflag[(int64_t)f * radix::kBins + ty0 + threadIdx.x]) s_robust = 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // (Staging the band's records in shared memory was measured: no gain -- the passes are issue-bound, not
    // load-latency bound -- and it costs a resident CTA, so the records are read from L2 in every pass.)
    const uint32_t n = e_hi - e_lo;
    const float4* __restrict__ it = items + frame_offsets[f];
    const uint32_t b_lo = e_lo, b_hi = e_hi;
    (void)stage_cap;

    // per-warp chunk, ends moved to run boundaries
    const uint32_t L = ((n + kBandThreads - 1) / kBandThreads) * 32;
    const uint32_t a_lo = run_start_at_or_after(it, min(b_lo + warp * L, b_hi), b_lo, b_hi, lane);
    const uint32_t a_hi = run_start_at_or_after(it, min(b_lo + (warp + 1) * L, b_hi), b_lo, b_hi, lane);
    __syncthreads();
    const bool robust = s_robust != 0;

    BandCtx c{s_acc, 0, 0, ty0, rows, C, TH, W, lane};
    for (int pass = 0; pass < npass; ++pass) {         // :33-34 xlim outer, ylim inner (npass = 4; fewer = profiling only)
        c.dx = pass >> 1;
        c.dy = pass & 1;
        if (robust) {
            sweep_two_pass<false>(c, it, a_lo, a_hi);
        } else {
            uint32_t pos = a_lo;
            float4 r = (pos + lane < a_hi) ? it[pos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            while (pos < a_hi) {
                const bool inb = pos + lane < a_hi;
                const uint32_t ck = inb ? run_key(r) : 0xffffffffu;
                const uint32_t pk = __shfl_up_sync(0xffffffffu, ck, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || ck != pk);
                // window = leading whole runs: cut at the last run head unless the chunk ends inside the window
                int take = (int)min(a_hi - pos, 32u);
                if (pos + 32 < a_hi) take = 31 - __clz(heads);
                if (take == 0) {
                    // a single run of >= 32 events: find its end, replay it with separate dt sweeps
                    const uint32_t k0 = __shfl_sync(0xffffffffu, ck, 0);
                    uint32_t end = pos + 32;
                    while (end < a_hi) {
                        const uint32_t j = end + lane;
                        const uint32_t k2 = (j < a_hi) ? run_key(it[j]) : 0xffffffffu;
                        const unsigned diff = __ballot_sync(0xffffffffu, k2 != k0);
                        if (diff) { end += (uint32_t)(__ffs(diff) - 1); break; }
                        end += 32;
                    }
                    end = min(end, a_hi);
                    sweep_two_pass<true>(c, it, pos, end);
                    pos = end;
                    r = (pos + lane < a_hi) ? it[pos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                    continue;
                }
                const uint32_t npos = pos + (uint32_t)take;
                const float4 rn = (npos + lane < a_hi) ? it[npos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);  // prefetch
                const bool on = lane < take;
                float w0, w1;
                const uint32_t a0 = corner(c, r, on, 0, &w0);
                const uint32_t a1 = corner(c, r, on, 1, &w1);
                commit_adjacent(c, a0, w0);            // :35 all dt = 0 adds of these runs ...
                commit_adjacent(c, a1, w1);            //     ... then their dt = 1 adds
                pos = npos;
                r = rn;
            }
        }
        __syncthreads();
    }
    // coalesced write-out of the band: out[f][c][ty0 + r][:]
    const int64_t HW = (int64_t)H * W;
    float* o = out + (int64_t)f * C * HW + (int64_t)ty0 * W;
    const int per_c = rows * W;
    for (int cc = 0; cc < C; ++cc) {
        const float* a = s_acc + (int64_t)cc * TH * W;
        float* oc = o + (int64_t)cc * HW;
        if (vec4 && ((reinterpret_cast<uintptr_t>(oc) & 15) == 0)) {
            const float4* a4 = reinterpret_cast<const float4*>(a);
            float4* o4 = reinterpret_cast<float4*>(oc);
            for (int i = threadIdx.x; i < per_c / 4; i += kBandThreads) __stcs(o4 + i, a4[i]);
        } else {
            for (int i = threadIdx.x; i < per_c; i += kBandThreads) __stcs(oc + i, a[i]);
        }
    }
}

}  // namespace tri
}  // namespace oess
