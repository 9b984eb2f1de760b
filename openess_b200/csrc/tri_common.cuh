// tri_common.cuh -- arithmetic shared by the trilinear voxeliser kernels (voxel_trilinear.cu, tri_band.cuh).
// Every float op mirrors one torch op of DSEC/dataset/representations.py:24-43 and rounds like it
// (explicit _rn intrinsics + --fmad=false: no FMA contraction).
#pragma once
#include "common.cuh"

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
// t_norm values whose .int() is INT_MIN on the reference's hosts (NaN / out of int32 range) can never pass the
// 0 <= tl < C test (representations.py:29,36): such events contribute nothing.
__device__ __forceinline__ bool t_reachable(float tn) { return tn > -2147483648.0f && tn < 2147483648.0f; }

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

}  // namespace tri
}  // namespace oess
