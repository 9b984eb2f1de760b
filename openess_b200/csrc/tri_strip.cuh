// tri_strip.cuh -- step 3 of the ORDERED trilinear voxeliser, strip formulation (default).
//
// Input: per frame, 16-byte records (x, y, t_norm, value) stably sorted by (source-cell row py = y0 + 1, source-cell
// column px = x0 + 1) by the row radix pass + k_rowsort, and coloff[f][py][2 s + j] = first record of row py whose
// column is >= min(s * WC, W) + j  (j = 0, 1; s = 0 .. NS), emitted by k_rowsort from its column scan.
//
// One WARP owns one output row Y of one strip of WC output columns [s WC, s WC + WC), all C time bins: C x WC
// accumulators in shared memory that no other warp ever touches, so there is no __syncthreads() anywhere and a
// slow (edge-clustered) strip delays nobody else.  The four (dx, dy) passes of representations.py:33-34 become four
// contiguous record ranges of the two source rows Y + 1 - dy:
//      dx = 0: columns px in [s WC + 1, s WC + WC + 1)        dx = 1: px in [s WC, s WC + WC)
// so x / y bounds hold by construction and every (event, dx, dy) triple is visited exactly once in the whole grid
// (the band formulation re-read and masked a 25 % halo).  LANES = EVENTS.  Lanes that hit the same accumulator (same
// cell, same time bin) are adjacent for time-sorted input: the group's head lane reads the accumulator once, folds
// the group's weights in event order out of its neighbours' registers (shfl_down) and writes it back once.
//   fast path   : both source rows hold <= 32 records for the strip (the common case: ~21 at 100 k events per
//                 frame) -> each row is loaded ONCE into registers and serves its dx = 0 and dx = 1 pass.
//   window path : longer ranges, 32 records at a time, cut at cell-run heads so that a cell's dt = 0 adds always
//                 precede its dt = 1 adds (:35); a single run of >= 32 events is replayed with separate dt sweeps.
//   robust path : rows flagged by k_rowsort (time bins not monotone inside a cell, i.e. input not time-sorted):
//                 separate dt sweeps, ranks from match_any.
#pragma once
#include "tri_band.cuh"

namespace oess {
namespace tri {

// One warp-wide slice of column-sorted records of one source row, with everything the (dx, dt) corners share.
struct Slice {
    float x, val;    // record fields still needed per dx
    float ay;        // 1 - |yl - y|                                  (representations.py:37)
    float at0, at1;  // 1 - |t0 - t_norm|, 1 - |t0 + 1 - t_norm|
    int x0, t0;
    bool on, ok0, ok1;   // lane holds a record; its dt = 0 / dt = 1 time bin is inside [0, C)
    // adjacency structure: lanes that hit the same accumulator (same cell, same time bin) are adjacent
    bool head;       // first lane of its group
    int cnt;         // heads: lanes in the group; others: 0
    int maxcnt;      // warp-uniform max group size (1 = no duplicates)
};

// px0 = x0 of the previous lane (shfl_up), already available from the window cut
__device__ __forceinline__ void slice_prep(Slice& S, const float4& r, bool on, int x0, int px0, int Y, int C, int lane) {
    S.on = on;
    S.x = r.x;
    S.val = r.w;
    S.x0 = x0;
    S.t0 = __float2int_rz(r.z);
    S.ay = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(Y), r.y)));
    S.at0 = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(S.t0), r.z)));
    S.at1 = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(S.t0 + 1), r.z)));
    S.ok0 = on && (unsigned)S.t0 < (unsigned)C;                                                   // :36
    S.ok1 = on && (unsigned)(S.t0 + 1) < (unsigned)C;
    const int pt0 = __shfl_up_sync(0xffffffffu, S.t0, 1);
    const bool same = on && lane > 0 && x0 == px0 && S.t0 == pt0;
    const unsigned dup = __ballot_sync(0xffffffffu, same);
    S.head = true;
    S.cnt = 1;
    S.maxcnt = 1;
    if (dup != 0) {
        const unsigned heads = ~dup;                                  // bit 0 is always a head
        const unsigned above = heads & ~((2u << lane) - 1u);          // heads strictly above this lane
        const int next = above ? (__ffs(above) - 1) : 32;
        S.head = !same;
        S.cnt = S.head ? next - lane : 0;
        S.maxcnt = __reduce_max_sync(0xffffffffu, S.cnt);
    }
}

// acc[a] = (((acc[a] + w_head) + w_head+1) + ...) in event order for every group with ok set (ok is uniform inside
// a group): the head lane pulls the group's weights out of its neighbours' registers.
template <bool DUP>
__device__ __forceinline__ void fold_commit(float* a, float w, bool ok, const Slice& S) {
    if (!DUP) {
        if (ok) *a = __fadd_rn(*a, w);                                                            // :43
    } else {
        const bool hd = ok && S.head;
        float v = 0.0f;
        if (hd) v = *a;
        v = __fadd_rn(v, w);
        float wk = __shfl_down_sync(0xffffffffu, w, 1);
        if (S.cnt > 1) v = __fadd_rn(v, wk);
        if (S.maxcnt > 2) {
#pragma unroll 4
            for (int k = 2; k < S.maxcnt; ++k) {
                wk = __shfl_down_sync(0xffffffffu, w, k);
                if (S.cnt > k) v = __fadd_rn(v, wk);
            }
        }
        if (hd) *a = v;
    }
    __syncwarp();
}

// the dt = 0 and dt = 1 corners of one (dx, dy) pass for a slice; `valid` = lane takes part in this pass
template <int WC, bool DUP>
__device__ __forceinline__ void slice_pass(float* acc, const Slice& S, int dx, int xbase, bool valid) {
    const int xl = S.x0 + dx;                                                                     // :33
    const float ax = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(xl), S.x)));
    const float pxy = __fmul_rn(__fmul_rn(S.val, ax), S.ay);                                       // :37 left to right
    float* a = acc + (S.t0 * WC + (xl - xbase));
    fold_commit<DUP>(a, __fmul_rn(pxy, S.at0), valid && S.ok0, S);                                 // :35 dt = 0 ...
    fold_commit<DUP>(a + WC, __fmul_rn(pxy, S.at1), valid && S.ok1, S);                            //     ... then dt = 1
}

struct StripCtx {
    float* acc;      // [C][WC] of this warp
    int dx, Y, xbase, C, lane;
};

// all events of [lo, hi): dt = 0 sweep, then dt = 1 sweep (any run length; ADJ = false: any time order)
template <int WC, bool ADJ>
__device__ __noinline__ void strip_sweeps(const StripCtx c, const float4* __restrict__ it, uint32_t lo, uint32_t hi) {
#pragma unroll 1
    for (int dt = 0; dt < 2; ++dt) {
#pragma unroll 1
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t i = base + c.lane;
            const bool on = i < hi;
            const float4 r = on ? it[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            const int x0 = on ? __float2int_rz(r.x) : (int)(0x80000000u | (unsigned)c.lane);
            const int t0 = __float2int_rz(r.z);
            const int xl = x0 + c.dx, tl = t0 + dt;
            const float w = weight_t(weight_xy(r.x, r.y, r.w, xl, c.Y), r.z, tl);
            const bool ok = on && (unsigned)tl < (unsigned)c.C;
            const uint32_t addr = (uint32_t)(tl * WC + (xl - c.xbase));
            if (ADJ) {
                Slice S;
                slice_prep(S, r, on, x0, __shfl_up_sync(0xffffffffu, x0, 1), c.Y, c.C, c.lane);
                fold_commit<true>(c.acc + addr, w, ok, S);
            } else {
                const unsigned peers = __match_any_sync(0xffffffffu, ok ? addr : (0x80000000u | (unsigned)c.lane));
                const int rank = __popc(peers & lanemask_lt());
                const int rounds = __reduce_max_sync(0xffffffffu, ok ? rank : 0) + 1;
                for (int k = 0; k < rounds; ++k) {
                    if (ok && rank == k) c.acc[addr] = __fadd_rn(c.acc[addr], w);
                    __syncwarp();
                }
            }
        }
    }
}

// one (dx, dy) pass over an arbitrary range, 32 records at a time (time-sorted rows)
template <int WC>
__device__ __forceinline__ void strip_windows(const StripCtx& c, const float4* __restrict__ it, uint32_t lo, uint32_t hi) {
    const int lane = c.lane;
    uint32_t pos = lo;
    float4 r = ((uint32_t)lane < hi - pos) ? it[pos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    while (pos < hi) {
        const uint32_t n = hi - pos;
        const bool inb = (uint32_t)lane < n;
        const int x0 = inb ? __float2int_rz(r.x) : (int)(0x80000000u | (unsigned)lane);
        const int px0 = __shfl_up_sync(0xffffffffu, x0, 1);
        int take = (int)n;
        if (n > 32u) {
            // not the last window: keep leading whole cell runs only
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || x0 != px0);
            take = 31 - __clz(heads);
            if (take == 0) {
                // one cell run of >= 32 events: find its end, replay it with separate dt sweeps
                const int k0 = __shfl_sync(0xffffffffu, x0, 0);
                uint32_t end = pos + 32;
                while (end < hi) {
                    const uint32_t j = end + lane;
                    const int k2 = (j < hi) ? __float2int_rz(it[j].x) : (int)0x80000000;
                    const unsigned diff = __ballot_sync(0xffffffffu, k2 != k0);
                    if (diff) { end += (uint32_t)(__ffs(diff) - 1); break; }
                    end += 32;
                }
                end = min(end, hi);
                strip_sweeps<WC, true>(c, it, pos, end);
                pos = end;
                r = ((uint32_t)lane < hi - pos && pos < hi) ? it[pos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
        }
        const uint32_t npos = pos + (uint32_t)take;
        const float4 rn = (npos + lane < hi) ? it[npos + lane] : make_float4(0.f, 0.f, 0.f, 0.f);   // prefetch
        const bool on = lane < take;
        Slice S;
        slice_prep(S, r, on, x0, px0, c.Y, c.C, lane);
        if (S.maxcnt == 1) slice_pass<WC, false>(c.acc, S, c.dx, c.xbase, on);
        else slice_pass<WC, true>(c.acc, S, c.dx, c.xbase, on);
        pos = npos;
        r = rn;
    }
}

// grid (ceil(NS / warps per CTA), H, F) or frame-major (F, H, ceil(NS / warps)); CT = compile-time C (0: runtime g.C)
template <int WC, int CT, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_strip_splat(const float4* __restrict__ items, const int64_t* __restrict__ frame_offsets,
              const uint32_t* __restrict__ coloff, const uint32_t* __restrict__ rowflag, Geom g, int NS,
              int F_, float* __restrict__ out) {
    extern __shared__ __align__(16) float s_strip[];  // [warps][C][WC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // frame-major launch order (OESS_STRIP_ORDER=1): consecutive CTAs take the same (row, strip) of consecutive frames,
    // so heavy (edge-clustered) and light frames interleave in time instead of forming a heavy tail
    const bool fmajor = gridDim.z != (unsigned)F_;
    const int s = (fmajor ? blockIdx.z : blockIdx.x) * (blockDim.x >> 5) + warp;
    if (s >= NS) return;                               // warps are independent: no CTA-wide barrier below
    const int Y = blockIdx.y, f = fmajor ? blockIdx.x : blockIdx.z;
    const int C = CT ? CT : g.C;
    const int H = g.H, W = g.W;
    const int xbase = s * WC;
    const int wlim = min(WC, W - xbase);
    const int KT = 2 * (NS + 1);
    float* acc = s_strip + warp * (C * WC);

    // range ends of the four passes (lanes 0..7: pass p = 2 dx + dy -> lanes 2p, 2p + 1) and the robust flags of
    // the two source rows (lanes 8, 9)
    uint32_t meta = 0;
    {
        const uint32_t* T = coloff + ((int64_t)f * (H + 1) + Y) * KT;      // row py = Y (dy = 1); py = Y + 1 is KT further
        if (lane < 8) {
            const int p = lane >> 1, end = lane & 1;
            meta = T[(1 - (p & 1)) * KT + 2 * (s + end) + 1 - (p >> 1)];
        } else if (lane < 10) {
            meta = rowflag[(int64_t)f * radix::kBins + Y + 9 - lane];
        }
    }
    if (CT) {
#pragma unroll
        for (int i = 0; i < (CT * WC / 4 + 31) / 32; ++i)
            if (i * 32 + lane < CT * WC / 4) reinterpret_cast<float4*>(acc)[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int i = lane; i < C * WC / 4; i += 32) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4* __restrict__ it = items + frame_offsets[f];
    // union ranges of the two source rows: A = row Y + 1 (dy = 0), B = row Y (dy = 1)
    const uint32_t loA = __shfl_sync(0xffffffffu, meta, 4), hiA = __shfl_sync(0xffffffffu, meta, 1);
    const uint32_t loB = __shfl_sync(0xffffffffu, meta, 6), hiB = __shfl_sync(0xffffffffu, meta, 3);
    const bool robust = (__shfl_sync(0xffffffffu, meta, 8) | __shfl_sync(0xffffffffu, meta, 9)) != 0;
    __syncwarp();                                      // representations.py:22 zeros visible

    if (!robust && hiA - loA <= 32u && hiB - loB <= 32u) {
        const bool onA = (uint32_t)lane < hiA - loA, onB = (uint32_t)lane < hiB - loB;
        const float4 rA = onA ? it[loA + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 rB = onB ? it[loB + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int xA = onA ? __float2int_rz(rA.x) : (int)(0x80000000u | (unsigned)lane);
        const int xB = onB ? __float2int_rz(rB.x) : (int)(0x80000000u | (unsigned)lane);
        Slice A, B;
        slice_prep(A, rA, onA, xA, __shfl_up_sync(0xffffffffu, xA, 1), Y, C, lane);
        slice_prep(B, rB, onB, xB, __shfl_up_sync(0xffffffffu, xB, 1), Y, C, lane);
        // the strip's columns for dx = 0 / dx = 1 (the union range holds one extra column at either end)
        const bool vA0 = onA && (unsigned)(xA - xbase) < (unsigned)wlim, vA1 = onA && (unsigned)(xA + 1 - xbase) < (unsigned)wlim;
        const bool vB0 = onB && (unsigned)(xB - xbase) < (unsigned)wlim, vB1 = onB && (unsigned)(xB + 1 - xbase) < (unsigned)wlim;
        if ((A.maxcnt | B.maxcnt) == 1) {              // :33-34 xlim outer, ylim inner
            slice_pass<WC, false>(acc, A, 0, xbase, vA0);
            slice_pass<WC, false>(acc, B, 0, xbase, vB0);
            slice_pass<WC, false>(acc, A, 1, xbase, vA1);
            slice_pass<WC, false>(acc, B, 1, xbase, vB1);
        } else {
            slice_pass<WC, true>(acc, A, 0, xbase, vA0);
            slice_pass<WC, true>(acc, B, 0, xbase, vB0);
            slice_pass<WC, true>(acc, A, 1, xbase, vA1);
            slice_pass<WC, true>(acc, B, 1, xbase, vB1);
        }
    } else {
        StripCtx c{acc, 0, Y, xbase, C, lane};
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            c.dx = pass >> 1;
            const uint32_t lo = __shfl_sync(0xffffffffu, meta, 2 * pass);
            const uint32_t hi = __shfl_sync(0xffffffffu, meta, 2 * pass + 1);
            if (lo >= hi) continue;
            if (__shfl_sync(0xffffffffu, meta, 8 + (pass & 1)) != 0) strip_sweeps<WC, false>(c, it, lo, hi);
            else strip_windows<WC>(c, it, lo, hi);
        }
    }
    __syncwarp();
    // write-out of the strip: out[f][c][Y][xbase .. xbase + wlim)
    const int64_t HW = (int64_t)H * W;
    float* o = out + (int64_t)f * C * HW + (int64_t)Y * W + xbase;
    if (wlim == WC && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(acc);
        if (CT) {
#pragma unroll
            for (int i = 0; i < (CT * WC / 4 + 31) / 32; ++i) {
                const int j = i * 32 + lane;
                if (j < CT * WC / 4) __stcs(reinterpret_cast<float4*>(o + (int64_t)(j / (WC / 4)) * HW) + (j % (WC / 4)), a4[j]);
            }
        } else {
            for (int j = lane; j < C * (WC / 4); j += 32)
                __stcs(reinterpret_cast<float4*>(o + (int64_t)(j / (WC / 4)) * HW) + (j % (WC / 4)), a4[j]);
        }
    } else {
        for (int cc = 0; cc < C; ++cc)
            for (int q = lane; q < wlim; q += 32) __stcs(o + (int64_t)cc * HW + q, acc[cc * WC + q]);
    }
}

}  // namespace tri
}  // namespace oess
