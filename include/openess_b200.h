/*
 * openess_b200.h -- C ABI of libopeness_b200.so: the B200-native (sm_100a) implementation of the
 * OpenESS per-step hot path.
 *
 * The reference (ldkong1205/OpenESS) is pure Python: it has no FFI and no operator registry, the
 * "plugin boundary" is a set of Python callables.  Every entry point below names the
 * reference callable it replaces (file:line into the reference tree); openess_b200/_lib.py is the
 * ctypes binding and INTEGRATION.md shows the stub a maintainer adds on the reference side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.
 *  - Every data pointer is a DEVICE pointer owned by the caller (e.g. the PyTorch allocator).
 *  - No hidden allocation: scratch space is a caller-provided workspace (`ws`, `ws_bytes`) whose size
 *    comes from the matching *_ws_bytes() query (host-side arithmetic only, no CUDA call).
 *  - No global mutable state, no host synchronisation: every call only enqueues work on `stream`
 *    and is safe to call concurrently from several host threads on different streams.
 *  - Return value: 0 = OK, <0 = argument error (OESS_E_*), >0 = cudaError_t of a failed launch.
 *  - Batched: one call voxelises F event-frames.  Events of all frames are concatenated;
 *    frame_offsets[F+1] (int64, device) delimits frame f as [frame_offsets[f], frame_offsets[f+1]).
 *    A frame with zero events yields an all-zero grid (the single-frame reference raises IndexError;
 *    the Python mirror raises the same before calling in).
 */
#ifndef OPENESS_B200_H
#define OPENESS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* oess_stream_t; /* == cudaStream_t */

#define OESS_OK 0
#define OESS_E_ARG (-1)       /* bad shape / null pointer / unsupported value            */
#define OESS_E_WORKSPACE (-2) /* workspace missing or too small                           */
#define OESS_E_RANGE (-3)     /* size exceeds what the kernels index (per-frame n >= 2^31) */

/* Accumulation mode of the voxelisers.
 *  ORDERED: replays the reference's sequential accumulation order per voxel (stable radix sort by
 *           pixel + per-pixel gather) -> bit-exact with the reference's single-thread CPU result,
 *           deterministic.
 *  ATOMIC : throughput mode, any accumulation order: result within 2e-5 + 1e-5 |v| of ORDERED (same noise class as the
 *           reference's own multi-threaded put_), not necessarily run-to-run deterministic.  Dense frames
 *           (> 200 000 events per frame on average) use float atomics (red.global.add.f32); on sparser frames the
 *           trilinear voxeliser runs the ORDERED pipeline, which is faster there than the L2 reduction rate allows the
 *           atomics kernel to be (OESS_ATOMIC_DISPATCH=0 forces the atomics kernel).            */
#define OESS_MODE_ORDERED 0
#define OESS_MODE_ATOMIC 1

/* Which voxeliser a workspace query is for. */
#define OESS_KIND_TRILINEAR 0
#define OESS_KIND_TBILINEAR 1

int oess_abi_version(void);
const char* oess_error_string(int code);

/* Measurement hooks (bench.py): number of kernels this library has launched so far in the process, and
 * an optional per-host-thread recorder that brackets every kernel with CUDA events on its stream.
 * oess_profile_end synchronises the recorded events and writes "kernel,launches,total_ms\n" lines. */
unsigned long long oess_launch_count(void);
int oess_profile_begin(void);
int oess_profile_end(char* buf, size_t buf_bytes);

/* Workspace size for oess_voxel_trilinear / oess_voxel_tbilinear_* (pure host arithmetic). */
int oess_voxel_ws_bytes(int kind, int mode, int64_t n_events_total, int n_frames, int C, int H, int W,
                        size_t* ws_bytes);

/* Replaces DSEC/dataset/representations.py:15-43 VoxelGrid.convert (trilinear x/y/t splat, float32).
 * x, y, pol, t: [n_events_total] float32 SoA exactly as the reference passes them (t already divided by
 * its last value by sequence_ov.py:155-156; convert() re-normalises with t[first], t[last] of the frame).
 * out: [F, C, H, W] float32, fully written (no pre-zeroing needed).
 * normalize != 0 additionally applies representations.py:45-53 per frame (nonzero mean / unbiased std). */
int oess_voxel_trilinear(const float* x, const float* y, const float* pol, const float* t,
                         const int64_t* frame_offsets, int64_t n_events_total, int n_frames, int C, int H,
                         int W, int mode, int normalize, float* out, void* ws, size_t ws_bytes,
                         oess_stream_t stream);

/* Replaces datasets/data_util.py:51-117 generate_voxel_grid (t-bilinear, integer pixels, float64 weights,
 * np.add.at order).  ev4: [n_events_total, 4] rows (x, y, t, p), int64 (DDD17 memmap path,
 * ddd17_events_loader.py:171-177) or float64 (DSEC non-voxel branch, sequence_ov.py:268-274).
 * mutate_p != 0 reproduces the reference's in-place `p[p == 0] = -1` on the caller's (device) array
 * (data_util.py:78-79).  out: [F, C, H, W] or, if separate_pol, [F, 2C, H, W] = concat(pos, neg). */
int oess_voxel_tbilinear_i64(int64_t* ev4, const int64_t* frame_offsets, int64_t n_events_total,
                             int n_frames, int C, int H, int W, int separate_pol, int mode, int mutate_p,
                             float* out, void* ws, size_t ws_bytes, oess_stream_t stream);
int oess_voxel_tbilinear_f64(double* ev4, const int64_t* frame_offsets, int64_t n_events_total,
                             int n_frames, int C, int H, int W, int separate_pol, int mode, int mutate_p,
                             float* out, void* ws, size_t ws_bytes, oess_stream_t stream);

/* Replaces datasets/data_util.py:17-35 generate_event_histogram -> out [F, 2, H, W] = stack(neg, pos).
 * Counts are exact integers in float32 (order independent), so there is a single mode.
 * status (device int32[1], may be NULL): set to 1 if any event indexes outside the H*W image (the
 * reference raises IndexError / wraps negative indices there; such events are skipped here). */
int oess_voxel_histogram_i64(int64_t* ev4, const int64_t* frame_offsets, int64_t n_events_total,
                             int n_frames, int H, int W, int mutate_p, float* out, int32_t* status,
                             oess_stream_t stream);
int oess_voxel_histogram_f64(double* ev4, const int64_t* frame_offsets, int64_t n_events_total,
                             int n_frames, int H, int W, int mutate_p, float* out, int32_t* status,
                             oess_stream_t stream);

/* The same two kernels on the DDD17 ON-DISK records (datasets/extract_data_tools/
 * example_loader_ddd17.py:32-38): t = `events.dat.t` int64 [n], xyp = `events.dat.xyp` int16 [n, 3] (x, y, p), staged
 * to the device as they are -- 14 B / event instead of the 32 B / event int64 rows extract_events_from_memmap (:41-54)
 * assembles on the host (np.concatenate + astype(int64) + column reorder).  Values are identical after widening, so
 * the outputs are bit-identical to oess_voxel_tbilinear_i64 / oess_voxel_histogram_i64 on the assembled rows.  The
 * reference's in-place p == 0 -> -1 lands on that temporary host copy, never on the memory map: nothing is written
 * back here.  Workspace: oess_voxel_ws_bytes(OESS_KIND_TBILINEAR, ...). */
int oess_voxel_tbilinear_ddd17(const int64_t* t, const int16_t* xyp, const int64_t* frame_offsets,
                               int64_t n_events_total, int n_frames, int C, int H, int W, int separate_pol, int mode,
                               float* out, void* ws, size_t ws_bytes, oess_stream_t stream);
int oess_voxel_histogram_ddd17(const int64_t* t, const int16_t* xyp, const int64_t* frame_offsets,
                               int64_t n_events_total, int n_frames, int H, int W, float* out, int32_t* status,
                               oess_stream_t stream);

/* Replaces DSEC/dataset/sequence_ov.py:204-210 rectify_events (gather rectify_map[y, x]) fused with the
 * per-chunk pre-step of sequence_ov.py:154-159 events_to_voxel_grid (t = f32(t - t[0]); t /= t[-1];
 * pol = f32(p)), per frame.  Raw DSEC records: x, y uint16, t int64 microseconds (t_offset already added,
 * DSEC/utils/eventslicer.py), p uint8.  rectify_map: [H, W, 2] float32.
 * status (device int32[1], may be NULL): 1 if any x >= W or y >= H (sequence_ov.py:208-209 asserts). */
int oess_dsec_rectify_tnorm(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p,
                            const float* rectify_map, const int64_t* frame_offsets, int64_t n_events_total,
                            int n_frames, int H, int W, float* xo, float* yo, float* po, float* to,
                            int32_t* status, oess_stream_t stream);
/* Same, for timestamps in the on-disk DSEC layout (uint32 microseconds relative to the file's t_offset,
 * DSEC/utils/eventslicer.py): only differences inside a frame are used, so the result is identical as long as
 * a frame does not wrap 2^32 us.  9 B/event of input instead of 13. */
int oess_dsec_rectify_tnorm_u32(const uint16_t* x, const uint16_t* y, const uint32_t* t, const uint8_t* p,
                                const float* rectify_map, const int64_t* frame_offsets, int64_t n_events_total,
                                int n_frames, int H, int W, float* xo, float* yo, float* po, float* to,
                                int32_t* status, oess_stream_t stream);

/* Replaces datasets/data_util.py:38-48 normalize_voxel_grid and e2vid/utils/inference_utils.py:77-85
 * (EventPreprocessor): x = (x != 0) * (x - mean) / std over the nonzero entries of each group.
 * x: [n_groups, group_numel] float32, in place.  stats: device float64 [n_groups, 3] = {sum, sumsq, nnz}.
 *   phase 0: compute stats and apply          (single-GPU / local-batch semantics)
 *   phase 1: compute stats only               (caller may all-reduce stats across ranks)
 *   phase 2: apply using the stats given
 * unbiased != 0 selects representations.py:45-53 semantics instead (torch.std, `std > 0` guard, only
 * nonzero entries are rewritten). */
int oess_nonzero_standardize(float* x, int64_t group_numel, int n_groups, double* stats, int phase,
                             int unbiased, oess_stream_t stream);

/* Replaces training/pretrain_trainer.py:445-465 (superpixel mean-pool via sparse one-hot matmul).
 * feat [B, Cf, H, W] f32 NCHW, seg [B, H, W] int64 superpixel ids; id' = id + b*S; ids outside [0, M)
 * are skipped and flagged in status.  pooled [M, Cf] = sum / (count + 1e-6), counts [M] f32.
 * The backward call scatters d_pooled back: d_feat[b, c, pix] = d_pooled[id', c] / (count[id'] + 1e-6). */
int oess_segpool_ws_bytes(int B, int Cf, int H, int W, int64_t M, size_t* ws_bytes);
int oess_segpool_fwd(const float* feat, const int64_t* seg, int B, int Cf, int H, int W, int S, int64_t M,
                     float* pooled, float* counts, int32_t* status, void* ws, size_t ws_bytes,
                     oess_stream_t stream);
int oess_segpool_bwd(const float* d_pooled, const int64_t* seg, const float* counts, int B, int Cf, int H,
                     int W, int S, int64_t M, float* d_feat, oess_stream_t stream);

/* Replaces utils/loss_functions.py:138-153 NCELoss: loss = mean_i CE((k q^T)/T, i).
 * k, q [M, D] f32.  loss: device f32[1].  dk/dq (may be NULL) receive d loss / d k, d q.
 * ws: M floats (row log-sum-exp) + M floats (column scratch); see oess_infonce_ws_bytes. */
int oess_infonce_ws_bytes(int64_t M, int D, size_t* ws_bytes);
int oess_infonce(const float* k, const float* q, int64_t M, int D, float temperature, float* loss,
                 float* dk, float* dq, void* ws, size_t ws_bytes, oess_stream_t stream);

/* Replaces utils/loss_functions.py:6-24 TaskLoss (= :96-135 DiceLoss + CrossEntropyLoss(ignore_index)).
 * logits [B, K, H, W] f32, target [B, H, W] int64.
 * partials: device float64 [2K+3] = {inter[K], denom[K], ce_sum, n_valid, n_bad}; exposed so that ranks can
 * all-reduce them for exact global-batch semantics.  n_bad counts targets that are neither a class in
 * [0, K) nor ignore_index (the reference raises on them in scatter_ / CrossEntropyLoss; the host checks the count).
 * A class equal to ignore_index contributes no Dice term but the sum is still divided by K (:127, :135).
 *   oess_dice_ce_partials: one pass over logits -> partials (accumulates into zeroed partials)
 *   oess_dice_ce_finish  : partials -> losses[3] = {dice, ce, w_dice*dice + w_ce*ce} (device f32)
 *   oess_dice_ce_bwd     : d(w_dice*dice + w_ce*ce)/d logits * grad_scale[0] -> d_logits            */
int oess_dice_ce_partials(const float* logits, const int64_t* target, int B, int K, int H, int W,
                          int64_t ignore_index, double* partials, oess_stream_t stream);
int oess_dice_ce_finish(const double* partials, int K, float w_dice, float w_ce, float* losses,
                        oess_stream_t stream);
int oess_dice_ce_finish_ex(const double* partials, int K, int64_t ignore_index, float w_dice, float w_ce, float* losses,
                           oess_stream_t stream);
int oess_dice_ce_bwd(const float* logits, const int64_t* target, int B, int K, int H, int W,
                     int64_t ignore_index, const double* partials, float w_dice, float w_ce,
                     const float* grad_scale, float* d_logits, oess_stream_t stream);

/* Replaces evaluation/metrics.py:4-23 semseg_compute_confusion: conf[gt, pred] += 1 over gt != ignore.
 * conf: device int64 [K, K], ACCUMULATES (zero it for a fresh matrix).  status: 1 if pred/gt out of range. */
int oess_confusion(const int64_t* pred, const int64_t* gt, int64_t n, int K, int64_t ignore_label,
                   int64_t* conf, int32_t* status, oess_stream_t stream);

/* Validation loop fused: training/base_trainer_ov.py:463-466 `pred = logits.argmax(dim=1)` +
 * evaluation/metrics.py:4-23 in one pass over logits [B, K, H, W] (K <= 64); conf ACCUMULATES like oess_confusion. */
int oess_argmax_confusion(const float* logits, const int64_t* gt, int B, int K, int H, int W, int64_t ignore_label,
                          int64_t* conf, int32_t* status, oess_stream_t stream);

/* Replaces the pointwise tail of e2vid/model/submodules.py:197-214 (ConvLSTM.forward after the Gates conv):
 * gates [B, 4C, H, W] = (in, remember, out, cell) chunks, prev_cell [B, C, H, W] or NULL (zero state):
 *   cell = sigmoid(remember) * prev_cell + sigmoid(in) * tanh(cell_gate);  hidden = sigmoid(out) * tanh(cell).
 * One fused pass (28 B / element) instead of ~10 elementwise kernels per encoder level and recurrent step. */
int oess_convlstm_gates(const float* gates, const float* prev_cell, float* hidden, float* cell, int B, int C,
                        int64_t HW, oess_stream_t stream);

/* Replaces training/openess_trainer.py:456 (:398, :497) torch.nn.L1Loss()(a, b) = mean |a - b|.
 * loss: device f32[1]; acc: device f64[1] scratch.  bwd: da = sign(a - b) * grad_scale[0] / n, db = -da. */
int oess_l1_mean(const float* a, const float* b, int64_t n, float* loss, double* acc, oess_stream_t stream);
int oess_l1_mean_bwd(const float* a, const float* b, int64_t n, const float* grad_scale, float* da, float* db,
                     oess_stream_t stream);

/* Replaces training/openess_trainer.py:460 (:402, :501) mean(1 - cosine_similarity(a, b, dim=1)) for
 * a, b [B, K, H, W] (eps = 1e-8, torch semantics).  loss: device f32[1]; acc: device f64[1] scratch. */
int oess_cos_consistency(const float* a, const float* b, int B, int K, int64_t HW, float* loss, double* acc,
                         oess_stream_t stream);
int oess_cos_consistency_bwd(const float* a, const float* b, int B, int K, int64_t HW, const float* grad_scale,
                             float* da, float* db, oess_stream_t stream);

/* Per-pixel linear map on NCHW tensors with Cin, Cout <= 64: y[b,k,p] = bias[k] + sum_c W[k,c] x[b,c,p].
 * Building block of the fused SemSegE2VID head: models/style_networks.py:163-165 chains conv1x1(32->256),
 * conv1x1(256->512) and conv(text_embeddings) with no non-linearity, so logits = (T W512 W256) x32 + T (W512 b256
 * + b512) -- computed directly, the 256/512-channel full-resolution maps (2.3 + 4.6 GB at batch 8) never exist.
 * W: [Cout, Cin] row-major; bias may be NULL.  Backward w.r.t. x = the same call with W^T and no bias.
 * wgrad: dW[k,c] = sum_{b,p} dy[b,k,p] x[b,c,p], db[k] = sum dy (db may be NULL); deterministic reduction. */
int oess_pixel_linear(const float* x, const float* W, const float* bias, int B, int Cin, int Cout, int64_t HW,
                      float* y, oess_stream_t stream);
int oess_pixel_linear_wgrad_ws_bytes(int Cin, int Cout, size_t* ws_bytes);
int oess_pixel_linear_wgrad(const float* dy, const float* x, int B, int Cin, int Cout, int64_t HW, float* dW,
                            float* db, void* ws, size_t ws_bytes, oess_stream_t stream);

/* Tensor-core GEMM (tcgen05.mma.kind::tf32, TMA-staged operands, accumulator in TMEM):
 *   C[M, N] = A[M, K] * B[N, K]^T + bias[N]     (A, B, C row-major float32; bias may be NULL)
 * = nn.Linear / a 1x1 convolution over channels-last pixels.  Replaces the cuBLAS / cuDNN call behind
 * models/image_model.py:121-124 (decoder conv 2048 -> 256), models/style_networks.py:163-165 (head convs) and the
 * ViT linears of models/maskclip_model.py:448-541.  Inputs are read as TF32 (what torch's default cuDNN path does
 * on the reference's GPU), accumulation is fp32; tolerance 2e-3 * sum_k |a||b|.
 * Requirements: K % 4 == 0 and 16-byte aligned pointers (TMA). */
int oess_gemm_tf32(const float* A, const float* B, const float* bias, float* C, int64_t M, int N, int K,
                   oess_stream_t stream);
/* Same GEMM with the epilogue the ViT blocks of models/maskclip_model.py:519-541 need:
 *   C = act(A * B^T + bias) + residual      act bit 0: GELU (erf form, mmcv FFN :507-513); bit 1: store C rounded to TF32
 *   (round-to-nearest) for outputs that are tensor-core operands next (q, k, v of the attention; see oess_conv2d_nhwc_tf32)
 * residual: [M, N] or NULL; may alias C (the `identity + dropout_layer(out)` residual of mmcv's MultiheadAttention / FFN). */
int oess_gemm_tf32_ex(const float* A, const float* B, const float* bias, const float* residual, float* C, int64_t M,
                      int N, int K, int act, oess_stream_t stream);

/* Training-time augmentation of a batch on the device (DSEC/dataset/sequence_ov.py:362-407).
 * oess_hflip_rows: in-place horizontal flip (torch.flip along the last dim) of the samples whose flag is non-zero.
 *   x: [B, rows_per_sample, W] of 4-byte (event tensor, frame) or 8-byte (label / pseudo-label / superpixel int64 maps)
 *   elements; flip: uint8 [B] on the device.
 * oess_frame_color_aug: frame [B, 3, H, W] float32 in [0, 1], in place, per sample b:
 *   frame = clamp(brightness[b] * frame, 0, 1)                       (TF.adjust_brightness; 1.0 = off)
 *   frame = clamp(contrast[b] * frame + (1 - contrast[b]) * mean(0.2989 r + 0.587 g + 0.114 b), 0, 1)   (TF.adjust_contrast)
 *   frame = frame + noise                                            (noise: [B, 3, H, W] or NULL; the caller draws it)
 *   gray_sums: float64 [B] scratch (zeroed here). */
int oess_hflip_rows(void* x, int elem_bytes, int B, int64_t rows_per_sample, int W, const uint8_t* flip,
                    oess_stream_t stream);
int oess_frame_color_aug(float* frame, int B, int64_t HW, const float* brightness, const float* contrast,
                         const float* noise, double* gray_sums, oess_stream_t stream);

/* Post-processing of reconstructed images (e2vid/image_reconstructor.py:126-140 PostProcessor =
 * e2vid/utils/inference_utils.py:234-252 UnsharpMaskFilter + :90-129 IntensityRescaler), img / out: [B, 1, H, W] float32
 * (out must not alias img), kernel5x5: the 25 taps of gkern(5, sigma) on the device:
 *   sharp = (1 + amount) * img - amount * conv2d(img, kernel5x5, padding = 2)          (amount <= 0: sharp = img)
 *   quantize != 0: out = float(uint8(clamp(255 * (sharp - imin) / (imax - imin), 0, 255))) / 255;   else out = sharp. */
int oess_unsharp_rescale(const float* img, const float* kernel5x5, int B, int H, int W, float amount, float imin,
                         float imax, int quantize, float* out, oess_stream_t stream);

/* Pooling layers of the torchvision-style ResNets (models/_resnet.py:137 MaxPool2d(kernel_size=3, stride=2, padding=1);
 * :149 AdaptiveAvgPool2d((1, 1))) on channels-last float32 tensors.  x: [B, H, W, C]; max pool y: [B, Ho, Wo, C] with
 * Ho = (H - 1) / 2 + 1 (C % 4 == 0); average pool y: [B, C] (mean over the HW pixels). */
int oess_maxpool3x3s2_nhwc(const float* x, int B, int H, int W, int C, float* y, oess_stream_t stream);
int oess_global_avgpool_nhwc(const float* x, int B, int64_t HW, int C, float* y, oess_stream_t stream);

/* E2VID decoder helpers (online reconstruction; e2vid/model/unet.py:165-168, submodules.py:34-63).
 * oess_zero_insert2x_nhwc: z[b, 2y, 2x, :] = x[b, y, x, :] + skip[b, y, x, :] (skip may be NULL), all other entries of the
 *   [B, 2H, 2W, C] output zero: ConvTranspose2d(k, stride 2, padding p, output_padding 1) of (x + skip) is then
 *   oess_conv2d_nhwc_tf32(z, rotated / transposed weights, stride 1, padding k - 1 - p).  C % 4 == 0.
 * oess_pred_sigmoid_nhwc: out[p] = sigmoid(dot(w, x[p, :] + skip[p, :]) + bias), the 1x1 prediction conv to one channel
 *   (eval BatchNorm folded into w / bias by the caller) + torch.sigmoid; C % 4 == 0, C <= 64; w 16-byte aligned. */
int oess_zero_insert2x_nhwc(const float* x, const float* skip, int B, int H, int W, int C, float* z, oess_stream_t stream);
/* F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False) of a contiguous plane tensor x [planes = B * C, h, w] and
 * its backward (models/deeplabv3.py:53-56 of the reference: full-resolution logits and 256-channel features).  The backward is
 * a separable gather (no atomics): tmp = planes * H * w floats of scratch. */
int oess_bilinear_resize_planes(const float* x, int64_t planes, int h, int w, int H, int W, float* out, oess_stream_t stream);
int oess_bilinear_resize_planes_bwd(const float* g, int64_t planes, int h, int w, int H, int W, float* tmp, float* dx,
                                    oess_stream_t stream);

/* Decoder transition of SemSegE2VID (models/style_networks.py:148-158): out [B, 2H, 2W, C1 + C2] = cat(nearest-neighbour x2
 * upsampling of x [B, H, W, C1], skip [B, 2H, 2W, C2]) along the channels, channels-last, one pass (C2 = 0 / skip = NULL: plain
 * upsampling).  _bwd: dx = 2 x 2 sums of g[..., :C1], dskip = g[..., C1:]; either output may be NULL.  C1, C2 % 4 == 0. */
int oess_upsample2x_cat_nhwc(const float* x, const float* skip, int B, int H, int W, int C1, int C2, float* out,
                             oess_stream_t stream);
int oess_upsample2x_cat_nhwc_bwd(const float* g, int B, int H, int W, int C1, int C2, float* dx, float* dskip,
                                 oess_stream_t stream);
int oess_pred_sigmoid_nhwc(const float* x, const float* skip, const float* w, float bias, int64_t pixels, int C,
                           float* out, oess_stream_t stream);

/* ---- MaskCLIP ViT-B/16 forward (models/maskclip_model.py) -- the non-GEMM kernels ------------------
 * Tokens are row-major [rows, D] float32 (rows = B * T, T = 1 + h * w); D % 128 == 0, D <= 1024.
 *
 * oess_vit_patchify: PatchEmbed input side (maskclip_model.py:427-441): 'corner' AdaptivePadding (zeros at the bottom /
 *   right up to a multiple of P) + non-overlapping P x P patches -> rows [B * h * w, C * P * P] in the column order
 *   (c, ky, kx) of projection.weight.view(D, -1), so that the strided conv is one oess_gemm_tf32 call.
 * oess_vit_assemble: x[b, 0] = cls + pos[0]; x[b, 1 + i] = tok[b, i] + pos[1 + i] (VisionTransformer.forward :799-806).
 * oess_layernorm_rows: y = LayerNorm(x) * gamma + beta over the last dim (biased variance, eps inside the sqrt).
 * oess_mha_fwd: softmax(Q K^T / sqrt(64)) V per (sample, head) on the packed in_proj output qkv [B, T, 3 * Hh * 64]
 *   (nn.MultiheadAttention layout: q | k | v, head-major inside each) -> out [B, T, Hh * 64]; head dim is 64 (ViT-B/16);
 *   fp32 online softmax, no attention matrix in memory.
 * oess_l2norm_rows: x /= ||x||_2 per row (MaskClipHead.cls_seg :218-219; no eps, as the reference).
 * oess_bilinear_tokens_to_nchw: bilinear resize (align_corners = False, mmseg.ops.resize = F.interpolate, :909-913) of a
 *   channels-last token map [B, h, w, K] to planes [B, K, H, W]. */
int oess_vit_patchify(const float* img, int B, int C, int H, int W, int P, float* rows, oess_stream_t stream);
int oess_vit_assemble(const float* tok, const float* cls, const float* pos, int B, int T, int D, float* x,
                      oess_stream_t stream);
int oess_layernorm_rows(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int D, float* y,
                        oess_stream_t stream);
int oess_mha_fwd(const float* qkv, int B, int T, int heads, float* out, oess_stream_t stream);
/* The same attention on the tensor cores (tc_mha.cu): both products as tcgen05.mma.kind::tf32 with S and the O tile in
 * TMEM, K / V tiles by TMA (V as an MN-major operand), fp32 online softmax in registers, P restaged through shared memory
 * as the K-major A operand.  TF32 operands / fp32 accumulate; same arguments and layout as oess_mha_fwd. */
int oess_mha_fwd_tc(const float* qkv, int B, int T, int heads, float* out, oess_stream_t stream);
int oess_l2norm_rows(float* x, int64_t rows, int D, oess_stream_t stream);
int oess_bilinear_tokens_to_nchw(const float* tok, int B, int h, int w, int K, int H, int W, float* out,
                                 oess_stream_t stream);

/* One ConvLSTM step of the E2VID recurrent encoder as one tensor-core kernel (tcgen05 implicit GEMM, TMA 4-D boxes
 * supply the 9 shifted taps with hardware zero padding, LSTM pointwise math fused into the TMEM epilogue).
 * Replaces e2vid/model/submodules.py:175-214 ConvLSTM.forward = cat(x, h_prev) -> Conv2d(2C, 4C, 3, padding=1) ->
 * chunk(in, remember, out, cell) -> c = sigmoid(remember) * c_prev + sigmoid(in) * tanh(cell); h = sigmoid(out) * tanh(c).
 * x, h_prev, c_prev, h_out, c_out: [B, H, W, C] CHANNELS-LAST float32; h_prev / c_prev may be NULL (zero state, :190-199).
 * w_packed: [4C, 2 * 9 * C], row chunk * 256 + gate * 64 + c (hidden channel chunk * 64 + c), column
 * (source: 0 = x, 1 = h; tap = ky * 3 + kx; channel) -- openess_b200/ops.py:convlstm_pack builds it from Gates.weight;
 * bias_packed: [4C] in the same row order.  C % 64 == 0.  TF32 operands, fp32 accumulate (tolerance 2e-3 * sum |x||w|). */
int oess_convlstm_step_nhwc(const float* x, const float* h_prev, const float* c_prev, const float* w_packed,
                            const float* bias_packed, float* h_out, float* c_out, int B, int H, int W, int C,
                            oess_stream_t stream);
/* bf16-operand variant for the FROZEN E2VID encoder (tcgen05.mma.kind::f16, fp32 accumulate): x, h_prev and w_packed are
 * bfloat16 with the layouts above; bias, cell state and the LSTM pointwise math stay fp32.  h_bf16_out (required): the hidden
 * state as the next step's operand; h_out (fp32, may be NULL): the same state for fp32 consumers (next encoder level, the
 * latent dictionary of e2vid/model/unet.py:163).  Stated tolerance vs the fp32 reference: see tests/test_tc_convlstm.py. */
int oess_convlstm_step_nhwc_bf16(const void* x, const void* h_prev, const float* c_prev, const void* w_packed,
                                 const float* bias_packed, float* h_out, void* h_bf16_out, float* c_out, int B, int H, int W,
                                 int C, oess_stream_t stream);
/* oess_conv2d_nhwc_tf32 (no residual) whose output is also (y != NULL) or only (y == NULL) stored as bfloat16. */
int oess_conv2d_nhwc_tf32_bf16out(const float* x, const float* w_packed, const float* bias, float* y, void* y_bf16, int B,
                                  int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil, int relu,
                                  oess_stream_t stream);

/* 2-D convolution over channels-last activations as a tcgen05 implicit GEMM (TF32 operands, fp32 accumulate), with
 * bias, optional residual add and optional ReLU fused into the TMEM epilogue.  Serves the frozen / inference
 * convolutions of the path: E2VID's strided 5x5 encoder convs with folded eval-mode BatchNorm + ReLU
 * (e2vid/model/submodules.py:7-31, e2vid/model/unet.py:128-135), 1x1 / 3x3 / dilated convs of models/_resnet.py:73-114
 * and models/deeplabv3.py:319-348.
 * x: [B, H, W, Cin] channels-last, Cin % 4 == 0; w_packed: [Cout, KH * KW * Cin_p] with Cin_p = Cin rounded up to 32
 * (zero padded), column (tap = ky * KW + kx, channel) -- openess_b200/ops.py:conv2d_pack; bias: [Cout] or NULL;
 * residual: [B, Ho, Wo, Cout] or NULL; y: [B, Ho, Wo, Cout], Ho = (H + 2 pad - dil (KH - 1) - 1) / stride + 1.
 * Zero padding is the TMA unit's out-of-bounds fill; stride is a strided TMA box traversal.
 * relu is a flag word: bit 0 = ReLU, bit 1 = store y rounded to TF32 (round-to-nearest, cvt.rna) -- for activations
 * that feed another tensor-core layer: the tensor core truncates fp32 operands to TF32, a systematic -2^-12 relative
 * bias per layer that accumulates when nothing re-normalises (eval-mode ResNets); rounding at the producer removes it. */
int oess_conv2d_nhwc_tf32(const float* x, const float* w_packed, const float* bias, const float* residual, float* y,
                          int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                          int relu, oess_stream_t stream);

/* oess_conv2d_nhwc_tf32 that additionally accumulates the BatchNorm batch statistics of its RAW output in the epilogue
 * (bn_sums[0..Cout) = sum over rows of y, bn_sums[Cout..2 Cout) = sum of y^2; device doubles, zeroed inside), and the
 * train-mode BatchNorm that consumes them (ws = the same buffer: its first 2 C doubles are bn_sums). */
int oess_conv2d_nhwc_tf32_stats(const float* x, const float* w_packed, const float* bias, float* y, int B, int H, int W,
                                int Cin, int Cout, int KH, int KW, int stride, int pad, int dil, double* bn_sums,
                                oess_stream_t stream);
int oess_batchnorm_nhwc_sums(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float eps, float momentum, const float* residual, int relu, void* ws,
                             size_t ws_bytes, oess_stream_t stream);

/* bfloat16-operand variants for the FROZEN networks of the path (teacher ResNet-50 models/image_model.py:32-62, E2VID encoder
 * convs e2vid/model/submodules.py:7-31): tcgen05.mma.kind::f16, fp32 accumulation, fp32 bias / residual / BatchNorm statistics.
 * x_bf16 [B, H, W, Cin] bf16 channels-last (Cin % 8 == 0); w_packed_bf16 [Cout, KH * KW * Cin_p] bf16, Cin_p = Cin rounded up to
 * 64; the result goes to y (fp32, may be NULL) and / or y_bf16 (may be NULL; Cout % 4 == 0).  bn_sums != NULL: batch statistics of
 * the raw output as oess_conv2d_nhwc_tf32_stats (then residual = NULL, relu = 0).  oess_batchnorm_nhwc_sums_bf16 is
 * oess_batchnorm_nhwc_sums whose result is also (write_f32 != 0) or only (write_f32 == 0) stored as bf16 in y_bf16. */
int oess_conv2d_nhwc_bf16(const void* x_bf16, const void* w_packed_bf16, const float* bias, const float* residual, float* y,
                          void* y_bf16, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int dil,
                          int relu, double* bn_sums, oess_stream_t stream);
int oess_batchnorm_nhwc_sums_bf16(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float eps, float momentum, const float* residual, int relu, void* y_bf16,
                                  int write_f32, void* ws, size_t ws_bytes, oess_stream_t stream);

/* Conv + InstanceNorm2d(affine=False) (+ residual) (+ ReLU) of the SemSegE2VID task decoder (models/style_networks.py:
 * 252-289 ReLUINSConv2d / INSResBlock) for its forward-only uses (validation, linear probing, test.py): the conv
 * accumulates PER-SAMPLE statistics in its epilogue (in_sums [B][2 Cout] doubles, zeroed inside), the second call
 * normalises y [B, HW, C] in place. */
int oess_conv2d_nhwc_tf32_instats(const float* x, const float* w_packed, const float* bias, float* y, int B, int H, int W,
                                  int Cin, int Cout, int KH, int KW, int stride, int pad, int dil, double* in_sums,
                                  oess_stream_t stream);
int oess_instancenorm_nhwc_sums(float* x, int B, int64_t HW, int C, const double* sums, float eps, const float* residual,
                                int relu, oess_stream_t stream);

/* Training variants of the conv + InstanceNorm block: the forward keeps x_hat (in place of the conv output) and writes
 * y = act(x_hat + residual) to y_out; the backward turns dy into the gradient dz of the conv output (and d_res):
 *   g = dy * [y > 0];  dz = inv_std * (g - mean_hw(g) - x_hat * mean_hw(g x_hat))      (y = NULL: no ReLU in the forward). */
int oess_instancenorm_nhwc_sums_train(float* x, int B, int64_t HW, int C, const double* sums, float eps,
                                      const float* residual, int relu, float* y_out, oess_stream_t stream);
int oess_instancenorm_nhwc_bwd(const float* dy, const float* y, const float* xhat, int B, int64_t HW, int C,
                               const double* fwd_sums, double* bwd_sums, float eps, float* dz, float* d_res,
                               oess_stream_t stream);

/* Training variants of conv -> BatchNorm (batch statistics from oess_conv2d_nhwc_tf32_stats in the first 2 C doubles of ws):
 * the forward keeps z (conv output) and writes y = act(BN(z) + residual) to y_out; the backward produces the gradient dz of
 * the conv output, d_res, and (d_beta, d_gamma) in bwd_sums [2 C] doubles.  models/deeplabv3.py head / ASPP blocks. */
int oess_batchnorm_nhwc_sums_train(float* z, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float eps, float momentum, const float* residual, int relu,
                                   float* y_out, void* ws, size_t ws_bytes, oess_stream_t stream);
int oess_batchnorm_nhwc_bwd(const float* dy, const float* y, const float* z, int64_t R, int C, const double* fwd_sums,
                            const float* gamma, double* bwd_sums, float eps, float* dz, float* d_res, oess_stream_t stream);

/* BatchNorm2d (torch.nn.BatchNorm2d semantics) over channels-last rows x [R = B*H*W, C], IN PLACE, with optional residual
 * add and ReLU: the normalisation between the teacher's tensor-core convolutions.  The OpenESS trainers call `.train()`
 * on the frozen ResNet-50 teacher every step (training/pretrain_trainer.py:370-371; models/image_model.py:116-117), so
 * training != 0 (batch statistics + running-statistics update with `momentum`, unbiased running variance) is the
 * reference's real operating point; training == 0 uses the running statistics.  No host synchronisation.
 * ws: oess_bn_ws_bytes(C) bytes of device scratch.  C % 4 == 0 and (C / 4 divides 256 or is a multiple of 256). */
int oess_bn_ws_bytes(int C, size_t* ws_bytes);
int oess_batchnorm_nhwc(float* x, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float eps, float momentum, int training, const float* residual, int relu,
                        void* ws, size_t ws_bytes, oess_stream_t stream);

/* Frame-branch teacher tail fused with the superpixel pooling that consumes it (never materialises the [B,256,H,W] map):
 *   pooled_sum[m] = sum_{pixels of superpixel m} normalize_C( bilinear upsample (align_corners=True) of d )[pixel]
 * Replaces models/image_model.py:121-124,139-141 (nn.Upsample x4 + F.normalize) followed by
 * training/pretrain_trainer.py:446-463 (one-hot sparse matmul; the division by count + 1e-6 stays with the caller).
 * d: [B, h, w, 256] channels-last decoder output; seg: int64 [B, H, W] ids (b * S added inside, ids outside [0, M) are
 * skipped and flagged in *status); pooled_sum [M, 256], counts [M]: zeroed and filled.  The backward entry takes the
 * gradient w.r.t. pooled_sum and produces d_grad [B, h, w, 256] (zeroed inside). */
int oess_upnorm_pool_fwd(const float* d, const int64_t* seg, int B, int h, int w, int C, int H, int W, int S, int64_t M,
                         float* pooled_sum, float* counts, int32_t* status, oess_stream_t stream);
int oess_upnorm_pool_bwd(const float* d, const int64_t* seg, const float* g_sum, int B, int h, int w, int C, int H, int W,
                         int S, int64_t M, float* d_grad, oess_stream_t stream);

/* x [B, C, HW] planes (NCHW) -> y [B, HW, Cp] channels-last, channels zero-padded to Cp (multiple of 4); stats != NULL
 * (double[3] = sum, sum of squares, non-zero count of x, from oess_nonzero_standardize phase 1) additionally applies the
 * EventPreprocessor normalisation of e2vid/utils/inference_utils.py:77-85 on the way.  Input transform in front of the
 * tensor-core head convolution of E2VID (e2vid/model/unet.py:126-127: 5 event channels). */
int oess_planes_to_nhwc_padded(const float* x, int B, int C, int64_t HW, const double* stats, int Cp, float* y,
                               oess_stream_t stream);
/* Same with every row zero-padded by pad_w pixels at both ends: y [B, H, W + 2 pad_w, Cp]. */
int oess_planes_to_nhwc_padded_w(const float* x, int B, int C, int H, int W, const double* stats, int Cp, int pad_w,
                                 float* y, oess_stream_t stream);
/* Thin-input stride-1 convolution (E2VID head conv, e2vid/model/unet.py:126-127: 5 -> 32 channels, 5 x 5): the KW taps of a
 * kernel row are folded into the channel dimension through an overlapping-window tensor map, K = KH * roundup(KW * Cin, 32)
 * instead of KH * KW * 32.  x: [B, H, W + KW - 1, Cin] channels-last, rows zero-padded by (KW - 1) / 2 pixels at both ends
 * (oess_planes_to_nhwc_padded_w), Cin % 4 == 0, KW * Cin <= 256; w_packed: [Cout, KH * roundup(KW * Cin, 32)] with column
 * (ky, kx * Cin + c); y: [B, H, W, Cout]; relu bit 0: ReLU, bit 1: store rounded to TF32.  TF32 operands, fp32 accumulate. */
int oess_conv2d_nhwc_tf32_rowunfold(const float* x, const float* w_packed, const float* bias, float* y, int B, int H, int W,
                                    int Cin, int Cout, int KH, int KW, int relu, oess_stream_t stream);

/* Weight gradient of a stride-1 convolution as a tcgen05 split-K GEMM over pixels (TF32 operands, fp32 accumulate):
 *   dW[co, ci, ky, kx] = sum_{b,y,x} dy[b, co, y, x] * x[b, ci, y + ky dil - pad, x + kx dil - pad]
 * x: [B, H, W, Cin], dy: [B, Ho, Wo, Cout] CHANNELS-LAST (both operands MN-major); dW: [Cout, Cin, KH, KW] (torch
 * layout), overwritten.  Cin % 4 == 0, Cout % 4 == 0, KW <= 5.
 * With oess_conv2d_nhwc_tf32 (forward, and backward-data on rotated weights) this is what torch autograd's
 * ConvolutionBackward does through cuDNN for the trainable convs of the path (models/image_model.py:121-124 decoder). */
int oess_conv2d_wgrad_nhwc_tf32(const float* x, const float* dy, float* dW, int B, int H, int W, int Cin, int Cout, int KH,
                                int KW, int pad, int dil, oess_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OPENESS_B200_H */
