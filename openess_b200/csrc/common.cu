// common.cu -- non-template helper kernels shared by the voxelisers + ABI bookkeeping.
#include "common.cuh"
#include "radix.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace oess {

__global__ void k_chunk_map(const int64_t* __restrict__ frame_offsets, int F, int chunk,
                            int* __restrict__ chunk_start) {
    // single CTA, 1024 threads: blocked exclusive scan of ceil(n_f / chunk)
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < F; base += blockDim.x) {
        const int f = base + tid;
        int v = 0;
        if (f < F) {
            const int64_t n = frame_offsets[f + 1] - frame_offsets[f];
            v = n > 0 ? (int)((n + chunk - 1) / chunk) : 0;
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            int x = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += u;
            }
            s_warp[lane] = x;  // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (w ? s_warp[w - 1] : 0) + incl - v;
        if (f < F) chunk_start[f] = excl;
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) chunk_start[F] = s_carry;
}

__global__ void __launch_bounds__(1024)
k_seg_exscan_u32(uint32_t* __restrict__ data, int64_t stride, int64_t len) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    uint32_t* seg = data + (int64_t)blockIdx.x * stride;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < len; base += 4096) {
        const int64_t i = base + (int64_t)tid * 4;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i + 3 < len) {
            v = *reinterpret_cast<const uint4*>(seg + i);
        } else {
            if (i < len) v.x = seg[i];
            if (i + 1 < len) v.y = seg[i + 1];
            if (i + 2 < len) v.z = seg[i + 2];
        }
        const uint32_t sum = v.x + v.y + v.z + v.w;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t x = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += u;
            }
            s_warp[lane] = x;
        }
        __syncthreads();
        const uint32_t excl = s_carry + (w ? s_warp[w - 1] : 0) + incl - sum;
        uint4 o4;
        o4.x = excl;
        o4.y = excl + v.x;
        o4.z = o4.y + v.y;
        o4.w = o4.z + v.z;
        if (i + 3 < len) {
            *reinterpret_cast<uint4*>(seg + i) = o4;
        } else {
            if (i < len) seg[i] = o4.x;
            if (i + 1 < len) seg[i + 1] = o4.y;
            if (i + 2 < len) seg[i + 2] = o4.z;
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + sum;
        __syncthreads();
    }
}

namespace radix {

__global__ void k_prefix_chunks(uint32_t* __restrict__ hist, const int* __restrict__ chunk_start,
                                uint32_t* __restrict__ tot, int nb) {
    const int f = blockIdx.x;
    const int b = blockIdx.y * 128 + threadIdx.x;
    if (b >= nb) {                                   // bins that cannot occur in this pass: never histogrammed
        tot[(int64_t)f * kBins + b] = 0;
        return;
    }
    const int c0 = chunk_start[f], c1 = chunk_start[f + 1];
    uint32_t run = 0;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {  // 4 independent loads in flight
        uint32_t* p = hist + (int64_t)c * kBins + b;
        const uint32_t v0 = p[0], v1 = p[kBins], v2 = p[2 * kBins], v3 = p[3 * kBins];
        p[0] = run;
        p[kBins] = run + v0;
        p[2 * kBins] = run + v0 + v1;
        p[3 * kBins] = run + v0 + v1 + v2;
        run += v0 + v1 + v2 + v3;
    }
    for (; c < c1; ++c) {
        uint32_t* p = hist + (int64_t)c * kBins + b;
        const uint32_t v = *p;
        *p = run;
        run += v;
    }
    tot[(int64_t)f * kBins + b] = run;
}

__global__ void __launch_bounds__(kBins) k_bin_scan(uint32_t* __restrict__ tot) {
    __shared__ uint32_t s_warp[32];
    uint32_t* t = tot + (int64_t)blockIdx.x * kBins;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t v = t[tid];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += u;
        }
        s_warp[lane] = x;
    }
    __syncthreads();
    t[tid] = (w ? s_warp[w - 1] : 0) + incl - v;
}

}  // namespace radix
}  // namespace oess

namespace oess {
namespace prof {

static std::atomic<unsigned long long> g_launches{0};

struct Rec {
    const char* name;
    cudaEvent_t e0, e1;
};
struct ThreadState {
    bool enabled = false;
    std::vector<Rec> recs;
};
static thread_local ThreadState t_state;

Scope::Scope(const char* n, cudaStream_t s) : name(n), st(s), rec(nullptr) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!t_state.enabled) return;
    Rec r{n, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, s);
    t_state.recs.push_back(r);
    rec = (void*)(uintptr_t)t_state.recs.size();   // index + 1
}
Scope::~Scope() {
    if (rec) cudaEventRecord(t_state.recs[(uintptr_t)rec - 1].e1, st);
}

}  // namespace prof
}  // namespace oess

OESS_API int oess_abi_version(void) { return 1; }

OESS_API unsigned long long oess_launch_count(void) {
    return oess::prof::g_launches.load(std::memory_order_relaxed);
}

OESS_API int oess_profile_begin(void) {
    auto& t = oess::prof::t_state;
    for (auto& r : t.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    t.recs.clear();
    t.enabled = true;
    return OESS_OK;
}

OESS_API int oess_profile_end(char* buf, size_t buf_bytes) {
    auto& t = oess::prof::t_state;
    t.enabled = false;
    std::map<std::string, std::pair<long long, double>> agg;
    std::vector<std::string> order;
    for (auto& r : t.recs) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        if (e != cudaSuccess) { t.recs.clear(); return (int)e; }
        auto it = agg.find(r.name);
        if (it == agg.end()) { order.push_back(r.name); agg[r.name] = {1, (double)ms}; }
        else { it->second.first += 1; it->second.second += ms; }
    }
    t.recs.clear();
    std::string out;
    char line[256];
    for (auto& n : order) {
        snprintf(line, sizeof(line), "%s,%lld,%.6f\n", n.c_str(), agg[n].first, agg[n].second);
        out += line;
    }
    if (!buf || buf_bytes < out.size() + 1) return OESS_E_WORKSPACE;
    memcpy(buf, out.c_str(), out.size() + 1);
    return OESS_OK;
}

OESS_API const char* oess_error_string(int code) {
    switch (code) {
        case OESS_OK: return "ok";
        case OESS_E_ARG: return "invalid argument";
        case OESS_E_WORKSPACE: return "workspace missing or too small";
        case OESS_E_RANGE: return "size exceeds kernel index range";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
