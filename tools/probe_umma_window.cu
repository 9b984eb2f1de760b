// Probe (tools only, not part of libopeness_b200.so): does a K-major SWIZZLE_128B tcgen05 shared-memory descriptor read a SHIFTED
// WINDOW of a larger TMA-written tile correctly -- start address not a multiple of 8 rows, 8-row groups 10 rows (1 280 B) apart?
// That is what a haloed A tile for the 3 x 3 convs needs.  One CTA: TMA loads A [180 rows, 32 tf32] and
// B [64 rows, 32 tf32] in the 128-byte swizzle, one M = 128, N = 64, K = 32 product with the A descriptor built from
// (row0, stride byte offset, base offset), D written to global memory.  The host side (probe_umma_window.py) compares with
// D[m][n] = sum_k A[row0 + (m / 8) * (sbo / 128) + m % 8][k] * B[n][k].
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -o tools/_probe/libprobe.so tools/probe_umma_window.cu -lcudart
#include <cuda_runtime.h>

#include "../openess_b200/csrc/tc_common.cuh"

using namespace oess::tc;

constexpr int kRowsA = 180, kRowsB = 64;

__global__ void __launch_bounds__(128, 1)
k_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int row0,
        int sbo_bytes, int base_off) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;                                   // 180 x 128 B (padded to 23 KB)
    uint8_t* sB = base + 23 * 1024;                       // 64 x 128 B
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + kRowsB * 128);
    uint64_t* done = full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(full, (kRowsA + kRowsB) * 128);
            tma_load_2d(sA, &tmA, full, 0, 0);
            tma_load_2d(sB, &tmB, full, 0, 0);
        }
        __syncwarp();
        mbar_wait(full, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)row0 * 128u;
            const uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
                                (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
            const uint64_t db = umma_desc_k128(smem_u32(sB));
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_tf32(tmem_acc, da + 2 * k, db + 2 * k, idesc, k != 0);
            umma_commit(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    float v[32];
    for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + ch * 32 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, const float* p, int rows) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
    const cuuint64_t dims[2] = {32, (cuuint64_t)rows}, strides[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)rows}, es[2] = {1, 1};
    return (int)((EncodeFn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// A: [180, 32] f32, B: [64, 32] f32, D: [128, 64] f32 (device pointers)
extern "C" int probe_umma_window(const float* A, const float* B, float* D, int row0, int sbo_bytes, int base_off) {
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, A, kRowsA);
    if (rc) return rc;
    rc = make_map(&tmB, B, kRowsB);
    if (rc) return rc;
    const int smem = 1024 + 23 * 1024 + kRowsB * 128 + 64;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(tmA, tmB, D, row0, sbo_bytes, base_off);
    return (int)cudaDeviceSynchronize();
}
