// augment.cu -- training-time augmentation of a DSEC batch on the device (DSEC/dataset/sequence_ov.py:
// 362-407): per-sample horizontal flip of the event tensor / frame / label / pseudo-label / superpixel maps, and the
// brightness / contrast / additive-noise chain on the frame (torchvision.transforms.functional.adjust_brightness /
// adjust_contrast semantics for float images).  HBM-bound: 8 B / element for a flipped sample (read + write in place),
// 2 passes over the frame (the contrast blend needs the mean grey level of the brightness-adjusted frame).
#include "common.cuh"

namespace oess {
namespace aug {

// One warp per row; rows of samples whose flag is 0 are skipped (no traffic).
template <class T>
__global__ void __launch_bounds__(256)
k_hflip_rows(T* __restrict__ x, int64_t rows_total, int64_t rows_per_sample, int W, const uint8_t* __restrict__ flip) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows_total; row += warps) {
        if (!flip[row / rows_per_sample]) continue;
        T* r = x + row * W;
        for (int i = lane; i < W / 2; i += 32) {
            const T a = r[i], b = r[W - 1 - i];
            r[i] = b;
            r[W - 1 - i] = a;
        }
    }
}

// pass 1: frame = clamp(bf * frame, 0, 1) in place; gsum[b] += sum over pixels of 0.2989 r + 0.587 g + 0.114 b (float64)
__global__ void __launch_bounds__(256)
k_brightness_gray(float* __restrict__ frame, int64_t HW, const float* __restrict__ brightness, double* __restrict__ gsum) {
    __shared__ double s_part[8];
    const int b = blockIdx.y;
    float* f = frame + (int64_t)b * 3 * HW;
    const float bf = brightness[b];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        float r = f[i], g = f[HW + i], bl = f[2 * HW + i];
        if (bf != 1.0f) {
            r = fminf(fmaxf(__fmul_rn(bf, r), 0.0f), 1.0f);
            g = fminf(fmaxf(__fmul_rn(bf, g), 0.0f), 1.0f);
            bl = fminf(fmaxf(__fmul_rn(bf, bl), 0.0f), 1.0f);
            f[i] = r; f[HW + i] = g; f[2 * HW + i] = bl;
        }
        acc += (double)__fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += s_part[k];
        atomicAdd(gsum + b, t);
    }
}

// pass 2: frame = clamp(cf * frame + (1 - cf) * mean, 0, 1) (+ noise)
__global__ void __launch_bounds__(256)
k_contrast_noise(float* __restrict__ frame, int64_t HW, const float* __restrict__ contrast, const double* __restrict__ gsum,
                 const float* __restrict__ noise) {
    const int b = blockIdx.y;
    const float cf = contrast[b];
    const float mean = (float)(gsum[b] / (double)HW);
    const float add = __fmul_rn(1.0f - cf, mean);
    float* f = frame + (int64_t)b * 3 * HW;
    const float* nz = noise ? noise + (int64_t)b * 3 * HW : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += (int64_t)gridDim.x * blockDim.x) {
        float v = f[i];
        if (cf != 1.0f) v = fminf(fmaxf(__fadd_rn(__fmul_rn(cf, v), add), 0.0f), 1.0f);
        if (nz) v = __fadd_rn(v, nz[i]);
        f[i] = v;
    }
}

// Post-processing of a reconstructed image (e2vid/utils/inference_utils.py:234-252 UnsharpMaskFilter, :90-129 IntensityRescaler):
//   sharp = (1 + amount) * img - amount * conv2d(img, gauss5x5, padding 2);
//   quantize: out = float(uint8(clamp(255 * (sharp - Imin) / (Imax - Imin), 0, 255))) / 255      else: out = sharp
__global__ void __launch_bounds__(256)
k_unsharp_rescale(const float* __restrict__ img, const float* __restrict__ kern, int H, int W, float amount, float imin,
                  float imax, int quantize, float* __restrict__ out) {
    __shared__ float s_k[25];
    if (threadIdx.x < 25) s_k[threadIdx.x] = kern[threadIdx.x];
    __syncthreads();
    const int64_t HW = (int64_t)H * W;
    const float* im = img + (int64_t)blockIdx.y * HW;
    float* o = out + (int64_t)blockIdx.y * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
        float blur = 0.0f;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                blur = __fadd_rn(blur, __fmul_rn(im[(int64_t)yy * W + xx], s_k[(dy + 2) * 5 + dx + 2]));
            }
        }
        float v = im[i];
        if (amount > 0.0f) v = __fsub_rn(__fmul_rn(1.0f + amount, v), __fmul_rn(amount, blur));
        if (quantize) {
            v = __fdiv_rn(__fmul_rn(255.0f, __fsub_rn(v, imin)), __fsub_rn(imax, imin));
            v = fminf(fmaxf(v, 0.0f), 255.0f);
            v = __fdiv_rn((float)(unsigned char)__float2int_rz(v), 255.0f);
        }
        o[i] = v;
    }
}

}  // namespace aug
}  // namespace oess

using namespace oess;

OESS_API int oess_unsharp_rescale(const float* img, const float* kernel5x5, int B, int H, int W, float amount, float imin,
                                  float imax, int quantize, float* out, oess_stream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || B > 65535) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!img || !kernel5x5 || !out || img == out) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = ((int64_t)H * W + 255) / 256;
    if (g > 4 * kNumSMs) g = 4 * kNumSMs;
    OESS_KERNEL("unsharp_rescale", st, aug::k_unsharp_rescale<<<dim3((unsigned)g, (unsigned)B), 256, 0, st>>>(
        img, kernel5x5, H, W, amount, imin, imax, quantize, out));
    return OESS_OK;
}

OESS_API int oess_hflip_rows(void* x, int elem_bytes, int B, int64_t rows_per_sample, int W, const uint8_t* flip,
                             oess_stream_t stream) {
    if (B < 0 || rows_per_sample <= 0 || W <= 0 || (elem_bytes != 4 && elem_bytes != 8)) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!x || !flip) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)B * rows_per_sample;
    int64_t g = (rows + 7) / 8;
    if (g > (int64_t)kNumSMs * 32) g = (int64_t)kNumSMs * 32;
    if (elem_bytes == 4)
        OESS_KERNEL("hflip_rows", st, aug::k_hflip_rows<uint32_t><<<(unsigned)g, 256, 0, st>>>((uint32_t*)x, rows, rows_per_sample, W, flip));
    else
        OESS_KERNEL("hflip_rows", st, aug::k_hflip_rows<unsigned long long><<<(unsigned)g, 256, 0, st>>>((unsigned long long*)x, rows, rows_per_sample, W, flip));
    return OESS_OK;
}

OESS_API int oess_frame_color_aug(float* frame, int B, int64_t HW, const float* brightness, const float* contrast,
                                  const float* noise, double* gray_sums, oess_stream_t stream) {
    if (B < 0 || HW <= 0 || B > 65535) return OESS_E_ARG;
    if (B == 0) return OESS_OK;
    if (!frame || !brightness || !contrast || !gray_sums) return OESS_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    OESS_CUDA(cudaMemsetAsync(gray_sums, 0, sizeof(double) * (size_t)B, st));
    int64_t g = (HW + 255) / 256;
    if (g > 2 * kNumSMs) g = 2 * kNumSMs;
    OESS_KERNEL("aug_brightness_gray", st, aug::k_brightness_gray<<<dim3((unsigned)g, (unsigned)B), 256, 0, st>>>(frame, HW, brightness, gray_sums));
    OESS_KERNEL("aug_contrast_noise", st, aug::k_contrast_noise<<<dim3((unsigned)g, (unsigned)B), 256, 0, st>>>(frame, HW, contrast, gray_sums, noise));
    return OESS_OK;
}
