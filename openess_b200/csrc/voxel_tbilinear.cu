// voxel_tbilinear.cu -- DDD17-style voxel grid (integer pixels, bilinear in t only) + 2-channel histogram.
// Replaces datasets/data_util.py:51-117 (generate_voxel_grid) and :17-35 (generate_event_histogram).
//
// Compile with --fmad=false.  Reference arithmetic: weights are float64, the
// grids are float32 and np.add.at performs, per event in ascending order, acc = f32(f64(acc) + w).
// Four passes: pos-left, pos-right, neg-left, neg-right into two grids; result pos - neg or concat.
//
// ORDERED mode (bit-exact): stable radix sort of {pixel key, event index} per frame (radix.cuh); then
// THREADS MAP TO SORTED EVENTS (dense lanes: pixels are mostly empty): the first event of every pixel run
// replays the run -- left pass then right pass, pos and neg accumulators in registers, float64 add +
// round-to-f32 per step exactly like np.add.at -- and writes that pixel's bins into the zeroed grid.
// Different pixels never interact (integer pixels), so there are no atomics and no conflicts.
// ATOMIC mode: one thread per event, 2 red.global.add.f32 (weights rounded to f32 first).
#include "common.cuh"
#include "radix.cuh"

namespace oess {
namespace tb {

struct Geom {
    int C, H, W;
    uint32_t invalid_key;  // H*W
};

template <class T> struct is_int { static constexpr bool value = false; };
template <> struct is_int<int64_t> { static constexpr bool value = true; };

struct Decoded {
    long long x, y, ti;
    double d, ap;
    bool pos, valid;
};

template <class T>
struct FrameTime {
    T first;
    double dT;
};

// Event rows as the reference holds them: an [N, 4] array (x, y, t, p) of int64 or float64 ...
template <class T>
struct RowsAoS {
    typedef T V;
    T* ev4;
    __device__ __forceinline__ V x(int64_t i) const { return ev4[i * 4 + 0]; }
    __device__ __forceinline__ V y(int64_t i) const { return ev4[i * 4 + 1]; }
    __device__ __forceinline__ V t(int64_t i) const { return ev4[i * 4 + 2]; }
    __device__ __forceinline__ V p(int64_t i) const { return ev4[i * 4 + 3]; }
    __device__ __forceinline__ void set_p(int64_t i, V v) const { ev4[i * 4 + 3] = v; }
    __host__ bool null() const { return ev4 == nullptr; }
};
// ... or the DDD17 on-disk records themselves (example_loader_ddd17.py:32-38: `events.dat.t` int64 [N, 1] and
// `events.dat.xyp` int16 [N, 3]), 14 B / event instead of the 32 B / event int64 rows that
// extract_events_from_memmap (:41-54) builds from them with np.concatenate(...).astype(int64)[:, [1, 2, 0, 3]]: the
// values are identical after the widening, so every kernel below computes the same bits.  The reference's in-place
// p == 0 -> -1 happens on that temporary copy, never on the memory map: set_p is a no-op.
struct RowsDDD17 {
    typedef int64_t V;
    const int64_t* ts;
    const int16_t* xyp;
    __device__ __forceinline__ V x(int64_t i) const { return (V)xyp[i * 3 + 0]; }
    __device__ __forceinline__ V y(int64_t i) const { return (V)xyp[i * 3 + 1]; }
    __device__ __forceinline__ V t(int64_t i) const { return ts[i]; }
    __device__ __forceinline__ V p(int64_t i) const { return (V)xyp[i * 3 + 2]; }
    __device__ __forceinline__ void set_p(int64_t, V) const {}
    __host__ bool null() const { return ts == nullptr || xyp == nullptr; }
};

template <class R>
__device__ __forceinline__ FrameTime<typename R::V> frame_time(const R& rows, int64_t fbeg, int64_t nf) {
    typedef typename R::V T;
    FrameTime<T> ft;
    ft.first = rows.t(fbeg);                               // data_util.py:68
    const T draw = rows.t(fbeg + nf - 1) - ft.first;       // :67,69
    ft.dT = (draw == (T)0) ? 1.0 : (double)draw;           // :71-72
    return ft;
}

template <class R>
__device__ __forceinline__ Decoded decode(const R& rows, int64_t i, const FrameTime<typename R::V>& ft, const Geom& g,
                                          bool mutate_p) {
    typedef typename R::V T;
    Decoded r;
    const T rx = rows.x(i), ry = rows.y(i), rt = rows.t(i);
    T p = rows.p(i);
    if (p == (T)0) { p = (T)-1; if (mutate_p) rows.set_p(i, p); }   // :78-79 (in place on the caller's array)
    const T num = (T)(g.C - 1) * (rt - ft.first);           // :76, array dtype arithmetic
    const double ts = __ddiv_rn((double)num, ft.dT);        //      then float64 true division
    r.x = is_int<T>::value ? (long long)rx : cvtt_f64_i64((double)rx);  // :74-75
    r.y = is_int<T>::value ? (long long)ry : cvtt_f64_i64((double)ry);
    r.ti = cvtt_f64_i64(ts);                                // :81
    r.d = __dsub_rn(ts, (double)r.ti);                      // :82
    r.ap = fabs((double)p);                                 // :83-84
    r.pos = (p == (T)1);                                    // :85
    r.valid = (r.x < g.W) && (r.x >= 0) && (r.y < g.H) && (r.y >= 0) && (ts >= 0.0) && (ts < (double)g.C);  // :88
    return r;
}

// The sort key of a row needs the frame's first/last timestamps (validity depends on ts), so k_keygen
// first turns rows into {key, index-in-frame} pairs; the radix passes then move 8-byte pairs only.
struct SrcPairs {
    typedef uint2 Item;
    const uint2* items;
    __device__ __forceinline__ Item load(int, int64_t fbeg, uint32_t li) const { return items[fbeg + li]; }
    __device__ __forceinline__ uint32_t key(const Item& it) const { return it.x; }
    __device__ __forceinline__ uint32_t key_at(int, int64_t fbeg, uint32_t li) const { return items[fbeg + li].x; }
};

template <class R>
__global__ void __launch_bounds__(256)
k_keygen(R rows, const int64_t* __restrict__ frame_offsets, const int* __restrict__ chunk_start,
         int F, Geom g, int mutate_p, uint2* __restrict__ pairs) {
    const int gch = blockIdx.x;
    const int f = find_frame(chunk_start, F, gch);
    if (f < 0) return;
    const int c = gch - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const int64_t nf = frame_offsets[f + 1] - fbeg;
    const FrameTime<typename R::V> ft = frame_time(rows, fbeg, nf);
#pragma unroll 2
    for (int s = 0; s < radix::kItemsPerThread; ++s) {
        const int64_t li = (int64_t)c * radix::kChunk + s * radix::kThreads + threadIdx.x;
        if (li >= nf) break;
        const Decoded d = decode(rows, fbeg + li, ft, g, mutate_p != 0);
        const uint32_t key = d.valid ? (uint32_t)(d.y * g.W + d.x) : g.invalid_key;
        pairs[fbeg + li] = make_uint2(key, (uint32_t)li);
    }
}

__device__ __forceinline__ float add_at(float acc, double w) {  // np.add.at(f32 grid, idx, f64 vals)
    return __double2float_rn(__dadd_rn((double)acc, w));
}

// One thread per sorted {key, index} pair; the head of each pixel run replays the run.
// CT > 0: all CT bins at once (2*CT register accumulators).  CT == 0: any C, one bin at a time.
template <class R, int CT>
__global__ void __launch_bounds__(256)
k_runs(R rows, const uint2* __restrict__ pairs, const int64_t* __restrict__ frame_offsets,
       const int* __restrict__ chunk_start, int F, Geom g, int separate_pol, float* __restrict__ out) {
    const int gch = blockIdx.x;
    const int f = find_frame(chunk_start, F, gch);
    if (f < 0) return;
    const int c = gch - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const int64_t nf = frame_offsets[f + 1] - fbeg;
    const FrameTime<typename R::V> ft = frame_time(rows, fbeg, nf);
    const int64_t HW = (int64_t)g.H * g.W;
    const int planes = separate_pol ? 2 * g.C : g.C;
    const uint2* pr = pairs + fbeg;
#pragma unroll 1
    for (int s = 0; s < radix::kItemsPerThread; ++s) {
        const int64_t i = (int64_t)c * radix::kChunk + s * radix::kThreads + threadIdx.x;
        if (i >= nf) break;
        const uint2 me = pr[i];
        if (me.x >= g.invalid_key) continue;                      // :88 invalid events sort last
        if (i > 0 && pr[i - 1].x == me.x) continue;               // not the head of its pixel run
        float* o = out + (int64_t)f * planes * HW + me.x;
        if (i + 1 >= nf || pr[i + 1].x != me.x) {
            // single-event pixel (the common case): one decode; acc = f32(f64(0) + w) = f32(w) per touched bin,
            // the other bins stay at the memset zero; pos - neg = +-f32(w) exactly (data_util.py:91-116)
            const Decoded d = decode(rows, fbeg + me.y, ft, g, false);
            const int64_t plane0 = (separate_pol && !d.pos) ? g.C : 0;
            const bool negate = !separate_pol && !d.pos;
            if (d.ti < g.C) {
                const float v = add_at(0.0f, __dmul_rn(d.ap, __dsub_rn(1.0, d.d)));
                o[(plane0 + d.ti) * HW] = negate ? __fsub_rn(0.0f, v) : v;
            }
            if (d.ti + 1 < g.C) {
                const float v = add_at(0.0f, __dmul_rn(d.ap, d.d));
                o[(plane0 + d.ti + 1) * HW] = negate ? __fsub_rn(0.0f, v) : v;
            }
            continue;
        }
        constexpr int NA = CT > 0 ? CT : 1;
        const int nbin_loops = CT > 0 ? 1 : g.C;
        for (int bin = 0; bin < nbin_loops; ++bin) {
            float ap[NA], an[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) { ap[k] = 0.0f; an[k] = 0.0f; }
#pragma unroll 1
            for (int right = 0; right < 2; ++right) {             // left pass (:91-92,:102-103) then right (:96-97,:107-108)
                for (int64_t j = i; j < nf; ++j) {
                    const uint2 pj = (j == i) ? me : pr[j];
                    if (pj.x != me.x) break;
                    const Decoded d = decode(rows, fbeg + pj.y, ft, g, false);
                    const long long tbin = d.ti + right;
                    if (!(tbin < g.C)) continue;                  // :87 / :94
                    if (CT == 0 && tbin != bin) continue;
                    const double w = right ? __dmul_rn(d.ap, d.d) : __dmul_rn(d.ap, __dsub_rn(1.0, d.d));
                    if (CT > 0) {
#pragma unroll
                        for (int k = 0; k < CT; ++k) {
                            const bool hit = (k == (int)tbin);
                            ap[k] = (hit && d.pos) ? add_at(ap[k], w) : ap[k];
                            an[k] = (hit && !d.pos) ? add_at(an[k], w) : an[k];
                        }
                    } else {
                        if (d.pos) ap[0] = add_at(ap[0], w); else an[0] = add_at(an[0], w);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                const int b = CT > 0 ? k : bin;
                if (separate_pol) {
                    o[(int64_t)b * HW] = ap[k];                             // :113-114 concat([pos, neg])
                    o[(int64_t)(g.C + b) * HW] = an[k];
                } else {
                    o[(int64_t)b * HW] = __fsub_rn(ap[k], an[k]);           // :116
                }
            }
        }
    }
}

template <class R>
__global__ void __launch_bounds__(256)
k_atomic(R rows, const int64_t* __restrict__ frame_offsets, const int* __restrict__ chunk_start,
         int F, Geom g, int separate_pol, int mutate_p, float* __restrict__ out) {
    const int gch = blockIdx.x;
    const int f = find_frame(chunk_start, F, gch);
    if (f < 0) return;
    const int c = gch - chunk_start[f];
    const int64_t fbeg = frame_offsets[f];
    const int64_t nf = frame_offsets[f + 1] - fbeg;
    const FrameTime<typename R::V> ft = frame_time(rows, fbeg, nf);
    const int64_t HW = (int64_t)g.H * g.W;
    const int planes = separate_pol ? 2 * g.C : g.C;
    float* o = out + (int64_t)f * planes * HW;
#pragma unroll 2
    for (int s = 0; s < radix::kItemsPerThread; ++s) {
        const int64_t li = (int64_t)c * radix::kChunk + s * radix::kThreads + threadIdx.x;
        if (li >= nf) break;
        const Decoded d = decode(rows, fbeg + li, ft, g, mutate_p != 0);
        if (!d.valid) continue;
        const int64_t pixel = d.y * g.W + d.x;
        const float sign = (separate_pol || d.pos) ? 1.0f : -1.0f;
        const int64_t plane0 = (separate_pol && !d.pos) ? g.C : 0;
        if (d.ti < g.C)
            atomicAdd(o + (plane0 + d.ti) * HW + pixel, sign * (float)(d.ap * (1.0 - d.d)));
        if (d.ti + 1 < g.C)
            atomicAdd(o + (plane0 + d.ti + 1) * HW + pixel, sign * (float)(d.ap * d.d));
    }
}

// data_util.py:17-35: out[f, 0] = neg counts, out[f, 1] = pos counts; flat index x + W*y.
// One thread per event over the flat event range; the frame is found by binary search in frame_offsets.
template <class R>
__global__ void __launch_bounds__(256)
k_histogram(R rows, const int64_t* __restrict__ frame_offsets, int64_t n, int F, int H, int W,
            int mutate_p, float* __restrict__ out, int32_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = F;  // frame_offsets[lo] <= i < frame_offsets[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (frame_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t HW = (int64_t)H * W;
    float* o = out + (int64_t)lo * 2 * HW;
    typedef typename R::V T;
    T p = rows.p(i);
    if (p == (T)0) { p = (T)-1; if (mutate_p) rows.set_p(i, p); }   // :26
    if (p != (T)1 && p != (T)-1) return;                       // :30-31 boolean masks
    const T rx = rows.x(i), ry = rows.y(i);
    const long long x = is_int<T>::value ? (long long)rx : cvtt_f64_i64((double)rx);  // :23-24
    const long long y = is_int<T>::value ? (long long)ry : cvtt_f64_i64((double)ry);
    const long long idx = x + (long long)W * y;
    if (idx < 0 || idx >= HW) { if (status) *status = 1; return; }
    atomicAdd(o + (p == (T)1 ? HW : 0) + idx, 1.0f);            // :33 stack([neg, pos])
}

struct Ws {
    int* chunk_start;
    uint32_t *hist, *tot;
    uint2 *a, *b;
    size_t bytes;
};

static Ws carve(void* ws, int mode, int64_t n, int F, int H, int W) {
    Ws r{};
    WsCarver c(ws);
    r.chunk_start = c.take<int>((size_t)F + 1);
    if (mode == OESS_MODE_ORDERED) {
        r.hist = c.take<uint32_t>((size_t)radix::max_chunks(n, F) * radix::kBins);
        r.tot = c.take<uint32_t>((size_t)F * radix::kBins);
        r.a = c.take<uint2>((size_t)n);
        r.b = c.take<uint2>((size_t)n);
    }
    r.bytes = c.total();
    return r;
}

template <class R>
static int run(R rows, const int64_t* frame_offsets, int64_t n, int F, int C, int H, int W, int separate_pol,
               int mode, int mutate_p, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n < 0 || F < 0 || C <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;   // data_util.py:59-62 asserts
    if (mode != OESS_MODE_ORDERED && mode != OESS_MODE_ATOMIC) return OESS_E_ARG;
    if ((int64_t)H * W + 2 >= (1ll << 31) || n >= (1ll << 31)) return OESS_E_RANGE;
    if (F == 0) return OESS_OK;
    if (!frame_offsets || !out || (n > 0 && rows.null())) return OESS_E_ARG;
    const Ws w = carve(ws, mode, n, F, H, W);
    if (!ws || ws_bytes < w.bytes) return OESS_E_WORKSPACE;
    const Geom g{C, H, W, (uint32_t)(H * W)};
    const int64_t HW = (int64_t)H * W;
    const int planes = separate_pol ? 2 * C : C;

    OESS_KERNEL("k_chunk_map", st, k_chunk_map<<<1, 1024, 0, st>>>(frame_offsets, F, radix::kChunk, w.chunk_start));
    const int64_t nch = radix::max_chunks(n, F);

    if (mode == OESS_MODE_ATOMIC) {
        OESS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)F * planes * HW, st));
        if (n > 0) {
            OESS_KERNEL("tb_atomic", st, k_atomic<R><<<(unsigned)nch, radix::kThreads, 0, st>>>(rows, frame_offsets, w.chunk_start, F, g,
                                                                  separate_pol, mutate_p, out));
        }
        return OESS_OK;
    }
    OESS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)F * planes * HW, st));
    if (n == 0) return OESS_OK;
    uint2* cur = w.a;
    OESS_KERNEL("tb_keygen", st, k_keygen<R><<<(unsigned)nch, radix::kThreads, 0, st>>>(
        rows, frame_offsets, w.chunk_start, F, g, mutate_p, w.a));
    const int bits = radix::key_bits(g.invalid_key + 1);
    const int passes = (bits + radix::kBits - 1) / radix::kBits;
    const int pbits = (bits + passes - 1) / passes;
    const uint32_t mask = (1u << pbits) - 1;
    for (int p = 0; p < passes; ++p) {
        uint2* dst = (cur == w.a) ? w.b : w.a;
        SrcPairs src{cur};
        int rc = radix::run_pass(src, frame_offsets, w.chunk_start, F, nch, p * pbits, mask, w.hist, w.tot,
                                 (uint32_t*)nullptr, 0, dst, st);
        if (rc) return rc;
        cur = dst;
    }
    if (C == 5) {
        OESS_KERNEL("tb_runs", st, k_runs<R, 5><<<(unsigned)nch, radix::kThreads, 0, st>>>(
            rows, cur, frame_offsets, w.chunk_start, F, g, separate_pol, out));
    } else {
        OESS_KERNEL("tb_runs", st, k_runs<R, 0><<<(unsigned)nch, radix::kThreads, 0, st>>>(
            rows, cur, frame_offsets, w.chunk_start, F, g, separate_pol, out));
    }
    return OESS_OK;
}

template <class R>
static int run_hist(R rows, const int64_t* frame_offsets, int64_t n, int F, int H, int W, int mutate_p,
                    float* out, int32_t* status, cudaStream_t st) {
    if (n < 0 || F < 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    if (F == 0) return OESS_OK;
    if (!frame_offsets || !out || (n > 0 && rows.null())) return OESS_E_ARG;
    OESS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)F * 2 * H * W, st));
    if (status) OESS_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (n == 0) return OESS_OK;
    OESS_KERNEL("k_histogram", st, k_histogram<R><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows, frame_offsets, n, F, H, W, mutate_p, out,
                                                               status));
    return OESS_OK;
}

}  // namespace tb
}  // namespace oess

using namespace oess;

int oess_voxel_tbilinear_ws_bytes_impl(int mode, int64_t n, int F, int C, int H, int W, size_t* out) {
    if (!out || n < 0 || F < 0 || C <= 0 || H <= 0 || W <= 0) return OESS_E_ARG;
    if (mode != OESS_MODE_ORDERED && mode != OESS_MODE_ATOMIC) return OESS_E_ARG;
    if ((int64_t)H * W + 2 >= (1ll << 31)) return OESS_E_RANGE;
    *out = tb::carve(nullptr, mode, n, F, H, W).bytes;
    return OESS_OK;
}

OESS_API int oess_voxel_tbilinear_i64(int64_t* ev4, const int64_t* frame_offsets, int64_t n, int F, int C, int H,
                                      int W, int separate_pol, int mode, int mutate_p, float* out, void* ws,
                                      size_t ws_bytes, oess_stream_t stream) {
    return tb::run(tb::RowsAoS<int64_t>{ev4}, frame_offsets, n, F, C, H, W, separate_pol, mode, mutate_p, out, ws, ws_bytes,
                            (cudaStream_t)stream);
}
OESS_API int oess_voxel_tbilinear_f64(double* ev4, const int64_t* frame_offsets, int64_t n, int F, int C, int H,
                                      int W, int separate_pol, int mode, int mutate_p, float* out, void* ws,
                                      size_t ws_bytes, oess_stream_t stream) {
    return tb::run(tb::RowsAoS<double>{ev4}, frame_offsets, n, F, C, H, W, separate_pol, mode, mutate_p, out, ws, ws_bytes,
                           (cudaStream_t)stream);
}

OESS_API int oess_voxel_histogram_i64(int64_t* ev4, const int64_t* frame_offsets, int64_t n, int F, int H, int W,
                                      int mutate_p, float* out, int32_t* status, oess_stream_t stream) {
    return tb::run_hist(tb::RowsAoS<int64_t>{ev4}, frame_offsets, n, F, H, W, mutate_p, out, status, (cudaStream_t)stream);
}
OESS_API int oess_voxel_histogram_f64(double* ev4, const int64_t* frame_offsets, int64_t n, int F, int H, int W,
                                      int mutate_p, float* out, int32_t* status, oess_stream_t stream) {
    return tb::run_hist(tb::RowsAoS<double>{ev4}, frame_offsets, n, F, H, W, mutate_p, out, status, (cudaStream_t)stream);
}

OESS_API int oess_voxel_tbilinear_ddd17(const int64_t* t, const int16_t* xyp, const int64_t* frame_offsets, int64_t n, int F,
                                        int C, int H, int W, int separate_pol, int mode, float* out, void* ws, size_t ws_bytes,
                                        oess_stream_t stream) {
    return tb::run(tb::RowsDDD17{t, xyp}, frame_offsets, n, F, C, H, W, separate_pol, mode, 0, out, ws, ws_bytes,
                   (cudaStream_t)stream);
}
OESS_API int oess_voxel_histogram_ddd17(const int64_t* t, const int16_t* xyp, const int64_t* frame_offsets, int64_t n, int F,
                                        int H, int W, float* out, int32_t* status, oess_stream_t stream) {
    return tb::run_hist(tb::RowsDDD17{t, xyp}, frame_offsets, n, F, H, W, 0, out, status, (cudaStream_t)stream);
}

int oess_voxel_trilinear_ws_bytes_impl(int mode, int64_t n, int F, int C, int H, int W, size_t* out);

OESS_API int oess_voxel_ws_bytes(int kind, int mode, int64_t n, int F, int C, int H, int W, size_t* ws_bytes) {
    if (kind == OESS_KIND_TRILINEAR) return oess_voxel_trilinear_ws_bytes_impl(mode, n, F, C, H, W, ws_bytes);
    if (kind == OESS_KIND_TBILINEAR) return oess_voxel_tbilinear_ws_bytes_impl(mode, n, F, C, H, W, ws_bytes);
    return OESS_E_ARG;
}
