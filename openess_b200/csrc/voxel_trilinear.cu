// voxel_trilinear.cu -- DSEC-style voxel grid: trilinear (x, y, t) splat of float-pixel events.
// Replaces DSEC/dataset/representations.py:15-55 (VoxelGrid.convert).
//
// Compile with --fmad=false: every float op below must round exactly like the reference's separate
// torch ops (no FMA contraction).  The rounding-critical expressions additionally use the explicit
// _rn intrinsics so the intent survives a flag change.
//
// ORDERED mode (bit-exact).  Per voxel the reference adds, in this order (representations.py:33-43,
// serial put_): for xl in (x0, x0+1): for yl in (y0, y0+1): for tl in (t0, t0+1): events ascending.
// A float sum is not associative, so that order is replayed exactly:
//   1. one stable radix pass (radix.cuh) sorts each frame's events by source-cell ROW (y0 + 1); while
//      loading, every event is turned into a 16-byte record (x, y, t_norm, 2*pol-1).  The pass's bin
//      scan doubles as the per-row CSR.
//   2. k_rowsort: one warp per (frame, row) stably counting-sorts its row by source-cell column in
//      shared memory -> records are now in (row, column, event order) = "cell-sorted".
//   3. k_band_splat: one CTA per band of TH output rows keeps the band's C x TH x W accumulators in
//      shared memory.  For each of the 4 (dx, dy) passes in reference order, THREADS MAP TO EVENTS
//      (dense lanes; pixels are 3x more numerous than events and mostly empty): the first event of every
//      cell run walks its run, first the dt=0 adds then the dt=1 adds, with plain shared-memory
//      read-add-write.  Inside one pass different cells hit different pixel columns, so there are no
//      conflicts and no atomics; __syncthreads() separates the passes.  The band is then written out
//      once, coalesced (no memset of the grid, no global atomics).
//   Frames whose sensor is taller than 1022 rows (or whose band does not fit shared memory) take the
//   generic path: two LSD radix passes over the full cell key + per-cell CSR + one thread per pixel.
// ATOMIC mode: one thread per event, 8 red.global.add.f32 into the zeroed grid.
#include "common.cuh"
#include "normalize.cuh"
#include "radix.cuh"

namespace oess {
namespace tri {

struct Geom {
    int C, H, W;
    uint32_t invalid_key;  // generic path: (H+1)*(W+1); events that cannot touch the grid sort last
};

// representations.py:27-28: source cell of an event.  px = x0 + 1 in [0, W], py = y0 + 1 in [0, H] are the
// cells that can reach the grid; anything else maps to the sentinel W + 1 / H + 1.
__device__ __forceinline__ uint32_t cell_px(float x, int W) {
    const int x0 = cvtt_f32_i32(x);
    return (x0 >= -1 && x0 <= W - 1) ? (uint32_t)(x0 + 1) : (uint32_t)(W + 1);
}
__device__ __forceinline__ uint32_t cell_py(float y, int H) {
    const int y0 = cvtt_f32_i32(y);
    return (y0 >= -1 && y0 <= H - 1) ? (uint32_t)(y0 + 1) : (uint32_t)(H + 1);
}
__device__ __forceinline__ uint32_t cell_key(float x, float y, const Geom& g) {  // generic path
    const uint32_t px = cell_px(x, g.W), py = cell_py(y, g.H);
    if (px > (uint32_t)g.W || py > (uint32_t)g.H) return g.invalid_key;
    return py * (uint32_t)(g.W + 1) + px;
}

// representations.py:24-25,31: per-event normalised time and polarity value, rounded like the reference.
__device__ __forceinline__ float t_norm(float t, float tfirst, float den, float cm1) {
    return __fdiv_rn(__fmul_rn(cm1, __fsub_rn(t, tfirst)), den);
}
__device__ __forceinline__ float pol_value(float pol) { return __fsub_rn(__fmul_rn(2.0f, pol), 1.0f); }

// representations.py:37: value * (1-|xl-x|) * (1-|yl-y|) * (1-|tl-t|), left to right.
__device__ __forceinline__ float weight_xy(float x, float y, float val, int xl, int yl) {
    const float ax = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(xl), x)));
    const float ay = __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(yl), y)));
    return __fmul_rn(__fmul_rn(val, ax), ay);
}
__device__ __forceinline__ float weight_t(float pxy, float tn, int tl) {
    return __fmul_rn(pxy, __fsub_rn(1.0f, fabsf(__fsub_rn(__int2float_rn(tl), tn))));
}

// record = (x, y, t_norm, value)
struct SrcSoA {
    typedef float4 Item;
    const float *x, *y, *pol, *t;
    const int64_t* frame_offsets;
    Geom g;
    int row_key;  // 1: key = cell row (banded path), 0: key = full cell key (generic path)
    __device__ __forceinline__ Item load(int f, int64_t fbeg, uint32_t li) const {
        const int64_t i = fbeg + li;
        const float tfirst = __ldg(t + fbeg);
        const float den = __fsub_rn(__ldg(t + frame_offsets[f + 1] - 1), tfirst);
        return make_float4(ld_stream(x + i), ld_stream(y + i),
                           t_norm(ld_stream(t + i), tfirst, den, (float)(g.C - 1)), pol_value(ld_stream(pol + i)));
    }
    __device__ __forceinline__ uint32_t key(const Item& it) const {
        return row_key ? cell_py(it.y, g.H) : cell_key(it.x, it.y, g);
    }
};
struct SrcAoS {
    typedef float4 Item;
    const float4* items;
    Geom g;
    __device__ __forceinline__ Item load(int, int64_t fbeg, uint32_t li) const { return items[fbeg + li]; }
    __device__ __forceinline__ uint32_t key(const Item& it) const { return cell_key(it.x, it.y, g); }
};

// ---------------------------------------------------------------------------------------------
// Banded path, step 2: stable counting sort of one (frame, row) segment by cell column.
// One warp per row; chunks of 32 events are ranked in order (match_any + popc), so ties keep
// event order.  rowoff = the radix pass's scanned bin totals ([F, 1024], bin = py).
// ---------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32)
k_rowsort(const float4* __restrict__ src, float4* __restrict__ dst, const int64_t* __restrict__ frame_offsets,
          const uint32_t* __restrict__ rowoff, int H, int W) {
    extern __shared__ uint32_t s_cnt_all[];           // [kRowWarps][W + 2]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + w;        // py in [0, H]
    const int f = blockIdx.y;
    if (row > H) return;
    const uint32_t* ro = rowoff + (int64_t)f * radix::kBins;
    const uint32_t s = ro[row], e = ro[row + 1];
    if (s == e) return;
    const int nb = W + 2;
    uint32_t* cnt = s_cnt_all + w * nb;
    const float4* in = src + frame_offsets[f];
    float4* out = dst + frame_offsets[f];
    for (int b = lane; b < nb; b += 32) cnt[b] = 0;
    __syncwarp();
    for (uint32_t i = s + lane; i < e; i += 32) atomicAdd(&cnt[cell_px(in[i].x, W)], 1u);
    __syncwarp();
    // exclusive scan of the W + 2 bins
    uint32_t carry = s;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const uint32_t v = (b < nb) ? cnt[b] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (b < nb) cnt[b] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    const unsigned lt = lanemask_lt();
    for (uint32_t i0 = s; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool act = i < e;
        float4 it = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t px = 0xffffffffu;
        if (act) { it = in[i]; px = cell_px(it.x, W); }
        const unsigned peers = __match_any_sync(0xffffffffu, px);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader && act) { base = cnt[px]; cnt[px] = base + __popc(peers); }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (act) out[base + __popc(peers & lt)] = it;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Banded path, step 3: ordered splat of one band of TH output rows from cell-sorted records.
// ---------------------------------------------------------------------------------------------
constexpr int kBandThreads = 256;

__device__ __forceinline__ uint32_t band_key(const float4& r, int H, int W) {
    return cell_py(r.y, H) * (uint32_t)(W + 2) + cell_px(r.x, W);
}

// One warp-wide slice of a (dx, dy, dt) pass: every lane in `lanes` owns one event.  Lanes that hit the same
// accumulator (same cell, same time bin) must add in event order: match_any groups them, the lane rank
// inside the group is the round in which the lane performs its plain shared-memory read-add-write.
__device__ __forceinline__ void splat_slice(float* __restrict__ acc, const float4& r, bool on, int dx, int dy,
                                            int dt, int ty0, int rows, int C, int TH, int H, int W, int lane) {
    uint32_t addr = 0x80000000u | (uint32_t)lane;      // unique sentinel: lane has nothing to add
    float wgt = 0.0f;
    if (on) {
        const int xl = (int)cell_px(r.x, W) - 1 + dx;   // x0 + dx        representations.py:33
        const int yl = (int)cell_py(r.y, H) - 1 + dy;   // y0 + dy        :34
        const int tl = (int)((unsigned)cvtt_f32_i32(r.z) + (unsigned)dt);   // :29,35
        const int yrel = yl - ty0;
        if (xl >= 0 && xl < W && yrel >= 0 && yrel < rows && tl >= 0 && tl < C) {   // :36 (or another band's row)
            wgt = weight_t(weight_xy(r.x, r.y, r.w, xl, yl), r.z, tl);              // :37
            addr = (uint32_t)((tl * TH + yrel) * W + xl);
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, addr);
    const int rank = __popc(peers & lanemask_lt());
    const bool valid = !(addr & 0x80000000u);
    const int rounds = __reduce_max_sync(0xffffffffu, valid ? rank + 1 : 0);
    for (int k = 0; k < rounds; ++k) {
        if (valid && rank == k) acc[addr] = __fadd_rn(acc[addr], wgt);              // :43 put_(accumulate=True)
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kBandThreads)
k_band_splat(const float4* __restrict__ items, const int64_t* __restrict__ frame_offsets,
             const uint32_t* __restrict__ rowoff, Geom g, int TH, float* __restrict__ out) {
    extern __shared__ __align__(16) float s_acc[];     // [C][TH][W]
    const int f = blockIdx.y;
    const int ty0 = blockIdx.x * TH;
    const int rows = min(TH, g.H - ty0);
    const int W = g.W, C = g.C, H = g.H;
    const int nacc = C * TH * W;
    for (int i = threadIdx.x; i < nacc; i += kBandThreads) s_acc[i] = 0.0f;   // :22 zeros

    // source-cell rows py in [ty0, ty0 + rows] feed output rows [ty0, ty0 + rows)
    const uint32_t* ro = rowoff + (int64_t)f * radix::kBins;
    const uint32_t e_lo = ro[ty0], e_hi = ro[min(ty0 + rows, H) + 1];
    const float4* it = items + frame_offsets[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();

    for (int pass = 0; pass < 4; ++pass) {             // :33-34 xlim outer, ylim inner
        const int dx = pass >> 1, dy = pass & 1;
        // Each warp takes 32-event windows; it OWNS the cell runs whose first event lies in its window and
        // follows its last run into the next windows, so a run is always replayed by one warp, in order.
        for (uint32_t base = e_lo + warp * 32; base < e_hi; base += kBandThreads) {
            const uint32_t i = base + lane;
            const bool act = i < e_hi;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t key = 0xffffffffu;
            if (act) { r = it[i]; key = band_key(r, H, W); }
            uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
            if (lane == 0) prev = (base > e_lo) ? band_key(it[base - 1], H, W) : 0xfffffffeu;
            const unsigned heads = __ballot_sync(0xffffffffu, act && key != prev);
            if (heads == 0) continue;                  // window lies inside a run owned by an earlier warp
            const int first = __ffs(heads) - 1;
            const uint32_t last_key = __shfl_sync(0xffffffffu, key, 31);
#pragma unroll 1
            for (int dt = 0; dt < 2; ++dt) {           // :35 tlim innermost: all dt=0 adds, then all dt=1 adds
                splat_slice(s_acc, r, act && lane >= first, dx, dy, dt, ty0, rows, C, TH, H, W, lane);
                uint32_t tail_key = last_key;          // follow the last run past the window
                for (uint32_t nb = base + 32; nb < e_hi; nb += 32) {
                    const uint32_t j = nb + lane;
                    const bool act2 = j < e_hi;
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    uint32_t key2 = 0xffffffffu;
                    if (act2) { q = it[j]; key2 = band_key(q, H, W); }
                    uint32_t prev2 = __shfl_up_sync(0xffffffffu, key2, 1);
                    if (lane == 0) prev2 = tail_key;
                    const unsigned heads2 = __ballot_sync(0xffffffffu, !act2 || key2 != prev2);
                    const int cont = heads2 ? __ffs(heads2) - 1 : 32;   // leading lanes continuing the run
                    if (cont == 0) break;
                    splat_slice(s_acc, q, lane < cont, dx, dy, dt, ty0, rows, C, TH, H, W, lane);
                    if (heads2) break;
                    tail_key = __shfl_sync(0xffffffffu, key2, 31);
                }
            }
        }
        __syncthreads();
    }
    // coalesced write-out of the band: out[f][c][ty0 + r][:]
    const int64_t HW = (int64_t)H * W;
    float* o = out + (int64_t)f * C * HW + (int64_t)ty0 * W;
    const int per_c = rows * W;
    for (int c = 0; c < C; ++c) {
        const float* a = s_acc + (int64_t)c * TH * W;
        float* oc = o + (int64_t)c * HW;
        for (int i = threadIdx.x; i < per_c; i += kBandThreads) __stcs(oc + i, a[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// Generic path: one thread per output pixel (CT accumulators) or voxel (CT == 0) walking the per-cell CSR.
// ---------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(256)
k_gather(const float4* __restrict__ items, const int64_t* __restrict__ frame_offsets,
         const uint32_t* __restrict__ pixoff, int64_t pix_stride, int F, Geom g, float* __restrict__ out) {
    const int64_t HW = (int64_t)g.H * g.W;
    const int64_t per_frame = CT > 0 ? HW : HW * g.C;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_frame * F) return;
    const int f = (int)(gid / per_frame);
    int64_t r = gid - (int64_t)f * per_frame;
    int my_t = 0;
    if (CT == 0) { my_t = (int)(r / HW); r -= (int64_t)my_t * HW; }
    const int yl = (int)(r / g.W), xl = (int)(r - (int64_t)yl * g.W);

    float acc[CT > 0 ? CT : 1];
#pragma unroll
    for (int c = 0; c < (CT > 0 ? CT : 1); ++c) acc[c] = 0.0f;

    const int64_t fbeg = frame_offsets[f];
    const int64_t nf = frame_offsets[f + 1] - fbeg;
    if (nf > 0) {
        const uint32_t* off = pixoff + (int64_t)f * pix_stride;
        const float4* fit = items + fbeg;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const uint32_t key = (uint32_t)(yl - dy + 1) * (uint32_t)(g.W + 1) + (uint32_t)(xl - dx + 1);
                const uint32_t s = off[key], e = off[key + 1];
                if (s == e) continue;
#pragma unroll
                for (int dt = 0; dt < 2; ++dt) {
                    for (uint32_t j = s; j < e; ++j) {
                        const float4 it = fit[j];
                        const int tl = (int)((unsigned)cvtt_f32_i32(it.z) + (unsigned)dt);
                        if (tl < 0 || tl >= g.C) continue;  // representations.py:36 (x/y already in range)
                        if (CT == 0 && tl != my_t) continue;
                        const float w = weight_t(weight_xy(it.x, it.y, it.w, xl, yl), it.z, tl);
                        if (CT > 0) {
#pragma unroll
                            for (int c = 0; c < CT; ++c) acc[c] = (c == tl) ? __fadd_rn(acc[c], w) : acc[c];
                        } else {
                            acc[0] = __fadd_rn(acc[0], w);
                        }
                    }
                }
            }
        }
    }
    if (CT > 0) {
        float* o = out + (int64_t)f * g.C * HW + (int64_t)yl * g.W + xl;
#pragma unroll
        for (int c = 0; c < CT; ++c) __stcs(o + (int64_t)c * HW, acc[c]);
    } else {
        __stcs(out + ((int64_t)f * g.C + my_t) * HW + (int64_t)yl * g.W + xl, acc[0]);
    }
}

__global__ void __launch_bounds__(256)
k_atomic(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ pol,
         const float* __restrict__ t, const int64_t* __restrict__ frame_offsets,
         const int* __restrict__ chunk_start, int F, Geom g, float* __restrict__ out) {
    const int gch = blockIdx.x;
    const int f = find_frame(chunk_start, F, gch);
    if (f < 0) return;
    const int c = gch - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const int64_t nf = frame_offsets[f + 1] - fbeg;
    const float tfirst = t[fbeg];
    const float den = __fsub_rn(t[fbeg + nf - 1], tfirst);
    const float cm1 = (float)(g.C - 1);
    const int64_t HW = (int64_t)g.H * g.W;
    float* o = out + (int64_t)f * g.C * HW;
#pragma unroll 4
    for (int s = 0; s < radix::kItemsPerThread; ++s) {
        const int64_t li = (int64_t)c * radix::kChunk + s * radix::kThreads + threadIdx.x;
        if (li >= nf) break;
        const int64_t i = fbeg + li;
        const float xi = ld_stream(x + i), yi = ld_stream(y + i), val = pol_value(ld_stream(pol + i));
        const float tn = t_norm(ld_stream(t + i), tfirst, den, cm1);
        const int x0 = cvtt_f32_i32(xi), y0 = cvtt_f32_i32(yi), t0 = cvtt_f32_i32(tn);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int xl = (int)((unsigned)x0 + (unsigned)(k >> 2));
            const int yl = (int)((unsigned)y0 + (unsigned)((k >> 1) & 1));
            const int tl = (int)((unsigned)t0 + (unsigned)(k & 1));
            if (xl < 0 || xl >= g.W || yl < 0 || yl >= g.H || tl < 0 || tl >= g.C) continue;
            atomicAdd(o + (int64_t)tl * HW + (int64_t)yl * g.W + xl, weight_t(weight_xy(xi, yi, val, xl, yl), tn, tl));
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct Plan {
    bool banded;
    int TH;
    size_t band_smem, row_smem;
};

static Plan make_plan(int C, int H, int W) {
    Plan p{};
    p.row_smem = sizeof(uint32_t) * (size_t)kRowWarps * (W + 2);
    // band height: largest power of two <= 8 whose accumulators fit ~100 KB (>= 2 CTAs / SM)
    int TH = 8;
    while (TH > 1 && sizeof(float) * (size_t)C * TH * W > 100 * 1024) TH >>= 1;
    p.TH = TH;
    p.band_smem = sizeof(float) * (size_t)C * TH * W;
    p.banded = (H + 2 <= radix::kBins) && p.band_smem <= 200 * 1024 && p.row_smem <= 200 * 1024;
    return p;
}

struct Ws {
    int* chunk_start;
    uint32_t *hist, *tot, *pix;
    float4 *a, *b;
    double* stats;
    int64_t pix_stride;
    size_t bytes;
};

static Ws carve(void* ws, int mode, int64_t n, int F, int C, int H, int W) {
    Ws r{};
    WsCarver c(ws);
    r.chunk_start = c.take<int>((size_t)F + 1);
    r.stats = c.take<double>((size_t)F * 3);
    if (mode == OESS_MODE_ORDERED) {
        r.hist = c.take<uint32_t>((size_t)radix::max_chunks(n, F) * radix::kBins);
        r.tot = c.take<uint32_t>((size_t)F * radix::kBins);
        if (!make_plan(C, H, W).banded) {
            const int64_t nkeys = (int64_t)(H + 1) * (W + 1) + 1;  // + invalid key
            r.pix_stride = (int64_t)align_up((size_t)nkeys + 1, 4);
            r.pix = c.take<uint32_t>((size_t)F * r.pix_stride);
        }
        r.a = c.take<float4>((size_t)n);
        r.b = c.take<float4>((size_t)n);
    }
    r.bytes = c.total();
    return r;
}

}  // namespace tri
}  // namespace oess

using namespace oess;

int oess_voxel_trilinear_ws_bytes_impl(int mode, int64_t n, int F, int C, int H, int W, size_t* out) {
    if (!out || n < 0 || F < 0 || C <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    if (mode != OESS_MODE_ORDERED && mode != OESS_MODE_ATOMIC) return OESS_E_ARG;
    if ((int64_t)(H + 1) * (W + 1) + 2 >= (1ll << 31)) return OESS_E_RANGE;
    *out = tri::carve(nullptr, mode, n, F, C, H, W).bytes;
    return OESS_OK;
}

OESS_API int oess_voxel_trilinear(const float* x, const float* y, const float* pol, const float* t,
                                  const int64_t* frame_offsets, int64_t n, int F, int C, int H, int W,
                                  int mode, int normalize, float* out, void* ws, size_t ws_bytes,
                                  oess_stream_t stream) {
    size_t need = 0;
    int rc = oess_voxel_trilinear_ws_bytes_impl(mode, n, F, C, H, W, &need);
    if (rc) return rc;
    if (F == 0) return OESS_OK;
    if (F > 65535) return OESS_E_RANGE;
    if (!frame_offsets || !out || (n > 0 && (!x || !y || !pol || !t))) return OESS_E_ARG;
    if (!ws || ws_bytes < need) return OESS_E_WORKSPACE;
    if (n >= (1ll << 31)) return OESS_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    const tri::Ws w = tri::carve(ws, mode, n, F, C, H, W);
    const int64_t HW = (int64_t)H * W;
    const int64_t nch = radix::max_chunks(n, F);

    if (mode == OESS_MODE_ATOMIC || n == 0) {
        OESS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)F * C * HW, st));
        if (n > 0) {
            const tri::Geom g{C, H, W, 0};
            OESS_KERNEL("k_chunk_map", st, k_chunk_map<<<1, 1024, 0, st>>>(frame_offsets, F, radix::kChunk, w.chunk_start));
            OESS_KERNEL("tri_atomic", st, tri::k_atomic<<<(unsigned)nch, radix::kThreads, 0, st>>>(
                x, y, pol, t, frame_offsets, w.chunk_start, F, g, out));
        }
    } else {
        const tri::Plan plan = tri::make_plan(C, H, W);
        OESS_KERNEL("k_chunk_map", st, k_chunk_map<<<1, 1024, 0, st>>>(frame_offsets, F, radix::kChunk, w.chunk_start));
        if (plan.banded) {
            const tri::Geom g{C, H, W, 0};
            tri::SrcSoA src{x, y, pol, t, frame_offsets, g, 1};
            rc = radix::run_pass(src, frame_offsets, w.chunk_start, F, nch, 0, radix::kBins - 1, w.hist, w.tot,
                                 (uint32_t*)nullptr, 0, w.a, st);
            if (rc) return rc;
            OESS_CUDA(cudaFuncSetAttribute(tri::k_rowsort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.row_smem));
            OESS_KERNEL("tri_rowsort", st, tri::k_rowsort<<<dim3((unsigned)((H + 1 + tri::kRowWarps - 1) / tri::kRowWarps), (unsigned)F),
                                                           tri::kRowWarps * 32, plan.row_smem, st>>>(
                w.a, w.b, frame_offsets, w.tot, H, W));
            OESS_CUDA(cudaFuncSetAttribute(tri::k_band_splat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.band_smem));
            OESS_KERNEL("tri_band_splat", st, tri::k_band_splat<<<dim3((unsigned)((H + plan.TH - 1) / plan.TH), (unsigned)F),
                                                                 tri::kBandThreads, plan.band_smem, st>>>(
                w.b, frame_offsets, w.tot, g, plan.TH, out));
        } else {
            const tri::Geom g{C, H, W, (uint32_t)((H + 1) * (W + 1))};
            OESS_CUDA(cudaMemsetAsync(w.pix, 0, sizeof(uint32_t) * (size_t)F * w.pix_stride, st));
            const int bits = radix::key_bits(g.invalid_key + 1);
            const int passes = (bits + radix::kBits - 1) / radix::kBits;
            const int pbits = (bits + passes - 1) / passes;
            const uint32_t mask = (1u << pbits) - 1;
            float4* cur = nullptr;
            for (int p = 0; p < passes; ++p) {
                float4* dst = (p & 1) ? w.b : w.a;
                if (p == 0) {
                    tri::SrcSoA src{x, y, pol, t, frame_offsets, g, 0};
                    rc = radix::run_pass(src, frame_offsets, w.chunk_start, F, nch, 0, mask, w.hist, w.tot, w.pix,
                                         w.pix_stride, dst, st);
                } else {
                    tri::SrcAoS src{cur, g};
                    rc = radix::run_pass(src, frame_offsets, w.chunk_start, F, nch, p * pbits, mask, w.hist, w.tot,
                                         (uint32_t*)nullptr, 0, dst, st);
                }
                if (rc) return rc;
                cur = dst;
            }
            OESS_KERNEL("k_seg_exscan_u32", st, k_seg_exscan_u32<<<(unsigned)F, 1024, 0, st>>>(
                w.pix, w.pix_stride, (int64_t)g.invalid_key + 2));
            if (C == 5) {
                const int64_t total = (int64_t)F * HW;
                OESS_KERNEL("tri_gather", st, tri::k_gather<5><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                    cur, frame_offsets, w.pix, w.pix_stride, F, g, out));
            } else {
                const int64_t total = (int64_t)F * HW * C;
                OESS_KERNEL("tri_gather", st, tri::k_gather<0><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                    cur, frame_offsets, w.pix, w.pix_stride, F, g, out));
            }
        }
    }
    if (normalize) {
        rc = launch_nonzero_standardize(out, (int64_t)C * HW, F, w.stats, 0, 1, st);
        if (rc) return rc;
    }
    return OESS_OK;
}
