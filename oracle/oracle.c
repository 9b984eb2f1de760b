/*
 * oracle.c -- CPU restatement of the OpenESS per-step hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for openess_b200.  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (openess_b200/) never imports, links or executes anything
 * under oracle/ and fails loudly when the CUDA library is missing.
 *
 * Every function restates, scalar by scalar, what the reference's numpy / torch-CPU code
 * computes, including its rounding behaviour and its quirks.
 * Citations are file:line into the reference tree (ldkong1205/OpenESS @ 5cb9f7f).
 *
 * Parity pin: the reference ships no tests / golden vectors for this path,
 * so the oracle is pinned against outputs of the reference's own Python code, generated in
 * the build container by oracle/make_golden.py and committed under tests/golden/
 * (tests/test_oracle_golden.py checks bit-equality for the voxelisers).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: no FMA contraction, plain SSE2
 * IEEE float/double arithmetic, which is what numpy and ATen's scalar kernels do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* x86 cvttss2si / cvttsd2si semantics: values that do not fit (and NaN) become INT_MIN.
 * This is what `tensor.int()` (representations.py:27-29) and `ndarray.astype(np.int64)`
 * (data_util.py:74-75,81) produce on the x86 hosts the reference runs on. */
static inline int32_t cvtt_f32_i32(float v) {
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT32_MIN;
    return (int32_t)v;
}
static inline int64_t cvtt_f64_i64(double v) {
    if (!(v > -9223372036854777856.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v;
}

/* ------------------------------------------------------------------------------------------
 * a2: datasets/data_util.py:51-117  generate_voxel_grid  (t-bilinear, integer pixels)
 *
 * ev:  [n,4] rows (x, y, t, p), C-contiguous; the polarity column is MUTATED (0 -> -1) exactly
 *      like the reference does on the caller's array (data_util.py:78-79).
 * out: [C,H,W] if !separate_pol, else [2C,H,W] = concat(pos, neg)  (data_util.py:113-117)
 * Accumulation is np.add.at(float32_grid, idx, float64_vals): sequential in event order, each
 * add rounds as f32(f64(acc) + w)  (SURVEY.md 0.5 / Appendix A.1).
 * returns 0, or -1 on bad arguments (the reference's asserts, data_util.py:59-62), -2 on n==0
 * (the reference raises IndexError on events[-1, 2]).
 * ---------------------------------------------------------------------------------------- */
#define TBILINEAR_BODY(T, IS_INT)                                                              \
    if (C <= 0 || H <= 0 || W <= 0) return -1;                                                 \
    if (n <= 0) return -2;                                                                     \
    const int64_t HW = (int64_t)H * W, CHW = HW * C;                                           \
    float* pos = (float*)calloc((size_t)CHW, sizeof(float));                                   \
    float* neg = (float*)calloc((size_t)CHW, sizeof(float));                                   \
    double* ts = (double*)malloc(sizeof(double) * (size_t)n);                                  \
    if (!pos || !neg || !ts) { free(pos); free(neg); free(ts); return -3; }                    \
    const T first = ev[2], last = ev[(n - 1) * 4 + 2];          /* data_util.py:67-68 */       \
    const T dTraw = last - first;                               /* :69 */                      \
    const double dT = (dTraw == 0) ? 1.0 : (double)dTraw;       /* :71-72 */                   \
    for (int64_t i = 0; i < n; ++i) {                                                          \
        /* :76  (C-1)*(t-first) in the array dtype, then true division in float64 */           \
        const T num = (T)(C - 1) * (ev[i * 4 + 2] - first);                                    \
        ts[i] = (double)num / dT;                                                              \
        if (ev[i * 4 + 3] == 0) ev[i * 4 + 3] = (T)-1;          /* :78-79 in-place */          \
    }                                                                                          \
    for (int pass = 0; pass < 4; ++pass) {                      /* :91-108 four np.add.at */   \
        const int want_pos = pass < 2, right = pass & 1;                                       \
        float* g = want_pos ? pos : neg;                                                       \
        for (int64_t i = 0; i < n; ++i) {                                                      \
            const int64_t x = IS_INT ? (int64_t)ev[i * 4] : cvtt_f64_i64((double)ev[i * 4]);   \
            const int64_t y = IS_INT ? (int64_t)ev[i * 4 + 1]                                  \
                                     : cvtt_f64_i64((double)ev[i * 4 + 1]);  /* :74-75 */      \
            const double t = ts[i];                                                            \
            const int64_t ti = cvtt_f64_i64(t);                 /* :81 */                      \
            const double d = t - (double)ti;                    /* :82 */                      \
            const T p = ev[i * 4 + 3];                                                         \
            const double ap = fabs((double)p);                  /* :83-84 np.abs(pols) */      \
            const int is_pos = (p == (T)1);                     /* :85 */                      \
            const int valid = (x < W) && (x >= 0) && (y < H) && (y >= 0) && (t >= 0) &&        \
                              (t < (double)C);                  /* :88 */                      \
            if (!valid || is_pos != want_pos) continue;                                        \
            const int64_t tb = ti + right;                                                     \
            if (!(tb < C)) continue;                            /* :87 / :94 */                \
            const double w = right ? ap * d : ap * (1.0 - d);                                  \
            const int64_t idx = x + y * W + tb * HW;                                           \
            g[idx] = (float)((double)g[idx] + w);               /* np.add.at f32 <- f64 */     \
        }                                                                                      \
    }                                                                                          \
    if (separate_pol) {                                                                        \
        memcpy(out, pos, sizeof(float) * (size_t)CHW);                                         \
        memcpy(out + CHW, neg, sizeof(float) * (size_t)CHW);    /* :113-114 */                 \
    } else {                                                                                   \
        for (int64_t i = 0; i < CHW; ++i) out[i] = pos[i] - neg[i];  /* :116 */                \
    }                                                                                          \
    free(pos); free(neg); free(ts);                                                            \
    return 0;

ORACLE_API int oracle_voxel_tbilinear_i64(int64_t* ev, int64_t n, int C, int H, int W,
                                          int separate_pol, float* out) {
    TBILINEAR_BODY(int64_t, 1)
}
ORACLE_API int oracle_voxel_tbilinear_f64(double* ev, int64_t n, int C, int H, int W,
                                          int separate_pol, float* out) {
    TBILINEAR_BODY(double, 0)
}

/* ------------------------------------------------------------------------------------------
 * a3: datasets/data_util.py:17-35  generate_event_histogram -> [2,H,W] = stack(neg, pos)
 * No bounds check in the reference (numpy raises IndexError / wraps negative indices);
 * the oracle reports out-of-range as error -4 (the product does the same).
 * ---------------------------------------------------------------------------------------- */
#define HISTOGRAM_BODY(T, IS_INT)                                                              \
    if (H <= 0 || W <= 0) return -1;                                                           \
    const int64_t HW = (int64_t)H * W;                                                         \
    memset(out, 0, sizeof(float) * (size_t)(2 * HW));                                          \
    for (int64_t i = 0; i < n; ++i)                                                            \
        if (ev[i * 4 + 3] == 0) ev[i * 4 + 3] = (T)-1;          /* :26 in-place */             \
    for (int64_t i = 0; i < n; ++i) {                                                          \
        const int64_t x = IS_INT ? (int64_t)ev[i * 4] : cvtt_f64_i64((double)ev[i * 4]);       \
        const int64_t y = IS_INT ? (int64_t)ev[i * 4 + 1] : cvtt_f64_i64((double)ev[i * 4 + 1]);\
        const T p = ev[i * 4 + 3];                                                             \
        if (p != (T)1 && p != (T)-1) continue;                  /* :30-31 masks */             \
        const int64_t idx = x + W * y;                                                         \
        if (idx < 0 || idx >= HW) return -4;                                                   \
        float* g = (p == (T)1) ? out + HW : out;                /* :33 stack([neg,pos]) */     \
        g[idx] = (float)((double)g[idx] + 1.0);                                                \
    }                                                                                          \
    return 0;

ORACLE_API int oracle_histogram_i64(int64_t* ev, int64_t n, int H, int W, float* out) {
    HISTOGRAM_BODY(int64_t, 1)
}
ORACLE_API int oracle_histogram_f64(double* ev, int64_t n, int H, int W, float* out) {
    HISTOGRAM_BODY(double, 0)
}

/* ------------------------------------------------------------------------------------------
 * a7: DSEC/dataset/representations.py:15-55  VoxelGrid.convert (trilinear splat, all float32)
 * x,y,pol,t: [n] f32.  out: [C,H,W] f32.  Serial put_(accumulate=True) order = pass-major,
 * event-minor (what a DataLoader worker with 1 intra-op thread does).
 * normalize != 0 applies representations.py:45-53 (nonzero mean / UNBIASED std) in float64
 * (tolerance-compared, torch's f32 reductions are order-dependent).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_voxel_trilinear(const float* x, const float* y, const float* pol,
                                      const float* t, int64_t n, int C, int H, int W,
                                      int normalize, float* out) {
    if (C <= 0 || H <= 0 || W <= 0) return -1;
    if (n <= 0) return -2;                                   /* t_norm[0] -> IndexError */
    const int64_t HW = (int64_t)H * W, CHW = HW * C;
    memset(out, 0, sizeof(float) * (size_t)CHW);             /* :22 clone of zeros */
    const float tfirst = t[0], tlast = t[n - 1];
    const float den = tlast - tfirst;                        /* :25 */
    const float cm1 = (float)(C - 1);
    for (int pass = 0; pass < 8; ++pass) {                   /* :33-35 nesting x, y, t */
        const int dx = (pass >> 2) & 1, dy = (pass >> 1) & 1, dt = pass & 1;
        for (int64_t i = 0; i < n; ++i) {
            const float tm = t[i] - tfirst;
            const float tn = (cm1 * tm) / den;               /* :25 */
            const int32_t x0 = cvtt_f32_i32(x[i]);           /* :27-29 */
            const int32_t y0 = cvtt_f32_i32(y[i]);
            const int32_t t0 = cvtt_f32_i32(tn);
            const int32_t xl = (int32_t)((uint32_t)x0 + (uint32_t)dx);
            const int32_t yl = (int32_t)((uint32_t)y0 + (uint32_t)dy);
            const int32_t tl = (int32_t)((uint32_t)t0 + (uint32_t)dt);
            if (!((xl < W) && (xl >= 0) && (yl < H) && (yl >= 0) && (tl >= 0) && (tl < C)))
                continue;                                    /* :36 */
            const float val = 2.0f * pol[i] - 1.0f;          /* :31 */
            const float ax = 1.0f - fabsf((float)xl - x[i]); /* :37 left-to-right products */
            const float ay = 1.0f - fabsf((float)yl - y[i]);
            const float at = 1.0f - fabsf((float)tl - tn);
            float w = val * ax;
            w = w * ay;
            w = w * at;
            const int64_t idx = HW * (int64_t)tl + (int64_t)W * yl + xl;   /* :39-41 */
            out[idx] = out[idx] + w;                         /* :43 put_ accumulate */
        }
    }
    if (normalize) {                                         /* :45-53 */
        double s = 0.0; int64_t nnz = 0;
        for (int64_t i = 0; i < CHW; ++i) if (out[i] != 0.0f) { s += out[i]; ++nnz; }
        if (nnz > 0) {
            const double mean = s / (double)nnz;
            double ss = 0.0;
            for (int64_t i = 0; i < CHW; ++i)
                if (out[i] != 0.0f) { const double d = out[i] - mean; ss += d * d; }
            const double std = sqrt(ss / (double)(nnz - 1)); /* torch.std: unbiased; nnz==1 -> NaN */
            for (int64_t i = 0; i < CHW; ++i) {
                if (out[i] == 0.0f) continue;
                out[i] = (std > 0) ? (float)((out[i] - (float)mean) / (float)std)
                                   : (float)(out[i] - (float)mean);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a5: DSEC/dataset/sequence_ov.py:204-210  rectify_events: (x', y') = rectify_map[y, x]
 * a6: DSEC/dataset/sequence_ov.py:154-159  events_to_voxel_grid pre-step:
 *       t = f32(t - t[0]);  t = t / t[-1];  pol = f32(p)
 * t is int64 microseconds (exact as float64 after the np.stack at sequence_ov.py:303).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_dsec_rectify_tnorm(const uint16_t* x, const uint16_t* y, const int64_t* t,
                                         const uint8_t* p, const float* rectify_map, int64_t n,
                                         int H, int W, float* xo, float* yo, float* po,
                                         float* to) {
    if (n <= 0) return -2;
    const int64_t t0 = t[0];
    const float tl = (float)(double)(t[n - 1] - t0);
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] >= W || y[i] >= H) return -4;               /* sequence_ov.py:208-209 asserts */
        const float* m = rectify_map + ((int64_t)y[i] * W + x[i]) * 2;
        xo[i] = m[0];
        yo[i] = m[1];
        po[i] = (float)p[i];
        const float tf = (float)(double)(t[i] - t0);         /* :155 */
        to[i] = tf / tl;                                     /* :156 (no guard: 0/0 -> NaN) */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a4/a8: datasets/data_util.py:38-48, e2vid/utils/inference_utils.py:77-85
 *   nnz = count(x != 0); mean = sum/nnz; std = sqrt(sumsq/nnz - mean^2) (biased, no eps);
 *   x = (x != 0) * (x - mean) / std.      Sums in float64 (tolerance-compared).
 * stats[3] = {sum, sumsq, nnz} returned for inspection.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_nonzero_standardize(float* x, int64_t n, double* stats) {
    double s = 0.0, ss = 0.0; int64_t nnz = 0;
    for (int64_t i = 0; i < n; ++i)
        if (x[i] != 0.0f) { s += x[i]; ss += (double)x[i] * x[i]; ++nnz; }
    if (stats) { stats[0] = s; stats[1] = ss; stats[2] = (double)nnz; }
    if (nnz > 0) {
        const float mean = (float)(s / (double)nnz);
        const float std = (float)sqrt(ss / (double)nnz - (s / (double)nnz) * (s / (double)nnz));
        for (int64_t i = 0; i < n; ++i) {
            const float m = (x[i] != 0.0f) ? 1.0f : 0.0f;
            x[i] = m * (x[i] - mean) / std;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a20: evaluation/metrics.py:4-23  confusion matrix, conf[gt, pred] += 1 over gt != ignore.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_confusion(const int64_t* pred, const int64_t* gt, int64_t n, int K,
                                int64_t ignore, int64_t* conf) {
    for (int64_t i = 0; i < n; ++i) {
        if (gt[i] == ignore) continue;
        const int64_t v = pred[i] + (int64_t)K * gt[i];       /* metrics.py:19 */
        if (v < 0 || v >= (int64_t)K * K) return -4;          /* bincount would grow -> assert :21 */
        conf[v] += 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a15: training/pretrain_trainer.py:445-465  superpixel mean-pool (sparse one-hot matmul).
 *   id' = id + b*S;  pooled[m, c] = sum_{pix: id'==m} feat[b, c, pix] / (count[m] + 1e-6)
 * feat: [B,Cf,H,W] f32 (NCHW), seg: [B,H,W] int64, pooled: [M,Cf], counts: [M] (float64 sums).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_segpool(const float* feat, const int64_t* seg, int B, int Cf, int H, int W,
                              int S, int64_t M, float* pooled, float* counts) {
    const int64_t HW = (int64_t)H * W;
    double* acc = (double*)calloc((size_t)(M * Cf), sizeof(double));
    double* cnt = (double*)calloc((size_t)M, sizeof(double));
    if (!acc || !cnt) { free(acc); free(cnt); return -3; }
    for (int b = 0; b < B; ++b)
        for (int64_t p = 0; p < HW; ++p) {
            const int64_t m = seg[b * HW + p] + (int64_t)b * S;
            if (m < 0 || m >= M) { free(acc); free(cnt); return -4; }
            cnt[m] += 1.0;
            for (int c = 0; c < Cf; ++c) acc[m * Cf + c] += feat[((int64_t)b * Cf + c) * HW + p];
        }
    for (int64_t m = 0; m < M; ++m) {
        const float den = (float)cnt[m] + 1e-6f;
        counts[m] = (float)cnt[m];
        for (int c = 0; c < Cf; ++c) pooled[m * Cf + c] = (float)acc[m * Cf + c] / den;
    }
    free(acc); free(cnt);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a16: utils/loss_functions.py:147-153  InfoNCE:  mean_i CE((k @ q^T)/T, i)   (float64)
 * Optional gradients dk, dq [M,D] of the scalar loss.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_infonce(const float* k, const float* q, int64_t M, int D, double temperature,
                              double* loss, double* dk, double* dq) {
    double* row = (double*)malloc(sizeof(double) * (size_t)M);
    if (!row) return -3;
    if (dk) memset(dk, 0, sizeof(double) * (size_t)(M * D));
    if (dq) memset(dq, 0, sizeof(double) * (size_t)(M * D));
    double total = 0.0;
    for (int64_t i = 0; i < M; ++i) {
        double mx = -INFINITY;
        for (int64_t j = 0; j < M; ++j) {
            double s = 0.0;
            for (int d = 0; d < D; ++d) s += (double)k[i * D + d] * q[j * D + d];
            row[j] = s / temperature;
            if (row[j] > mx) mx = row[j];
        }
        double z = 0.0;
        for (int64_t j = 0; j < M; ++j) z += exp(row[j] - mx);
        total += (mx + log(z)) - row[i];
        if (dk || dq)
            for (int64_t j = 0; j < M; ++j) {
                const double g = (exp(row[j] - mx) / z - (i == j ? 1.0 : 0.0)) /
                                 ((double)M * temperature);
                for (int d = 0; d < D; ++d) {
                    if (dk) dk[i * D + d] += g * q[j * D + d];
                    if (dq) dq[j * D + d] += g * k[i * D + d];
                }
            }
    }
    *loss = total / (double)M;
    free(row);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a17: utils/loss_functions.py:17-24 (TaskLoss), :114-135 (DiceLoss), :80-90 (BinaryDiceLoss)
 *   CE(ignore) mean over valid pixels  +  mean_c [ 1 - (2*sum(p_c*t_c) + 1)/(sum(p_c^2 + t_c^2) + 1) ]
 *   with p = softmax(logits) * mask, t = onehot(target*mask) * mask, sums over batch and pixels.
 * logits [B,K,H,W] f32, target [B,H,W] int64.  out[0]=dice, out[1]=ce, out[2]=dice+ce (float64).
 * dlogits (optional, [B,K,H,W] float64) = d(dice*w_dice + ce*w_ce)/dlogits.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_dice_ce(const float* logits, const int64_t* target, int B, int K, int H, int W,
                              int64_t ignore, double w_dice, double w_ce, double* out,
                              double* dlogits) {
    const int64_t HW = (int64_t)H * W;
    double* inter = (double*)calloc((size_t)K, sizeof(double));
    double* den = (double*)calloc((size_t)K, sizeof(double));
    double* p = (double*)malloc(sizeof(double) * (size_t)K);
    if (!inter || !den || !p) return -3;
    double ce = 0.0; int64_t nvalid = 0;
    for (int b = 0; b < B; ++b)
        for (int64_t px = 0; px < HW; ++px) {
            const int64_t tg = target[b * HW + px];
            if (tg == ignore) continue;
            if (tg < 0 || tg >= K) return -4;
            double mx = -INFINITY;
            for (int c = 0; c < K; ++c) {
                p[c] = logits[((int64_t)b * K + c) * HW + px];
                if (p[c] > mx) mx = p[c];
            }
            double z = 0.0;
            for (int c = 0; c < K; ++c) z += exp(p[c] - mx);
            ce += (mx + log(z)) - p[tg];
            ++nvalid;
            for (int c = 0; c < K; ++c) {
                const double pc = exp(p[c] - mx) / z;
                const double tc = (c == tg) ? 1.0 : 0.0;
                inter[c] += pc * tc;
                den[c] += pc * pc + tc * tc;
            }
        }
    double dice = 0.0;
    for (int c = 0; c < K; ++c) dice += 1.0 - (2.0 * inter[c] + 1.0) / (den[c] + 1.0);
    dice /= (double)K;
    const double cem = ce / (double)nvalid;                   /* 0/0 -> NaN like torch */
    out[0] = dice; out[1] = cem; out[2] = dice + cem;
    if (dlogits) {
        memset(dlogits, 0, sizeof(double) * (size_t)((int64_t)B * K * HW));
        for (int b = 0; b < B; ++b)
            for (int64_t px = 0; px < HW; ++px) {
                const int64_t tg = target[b * HW + px];
                if (tg == ignore) continue;
                double mx = -INFINITY;
                for (int c = 0; c < K; ++c) {
                    p[c] = logits[((int64_t)b * K + c) * HW + px];
                    if (p[c] > mx) mx = p[c];
                }
                double z = 0.0;
                for (int c = 0; c < K; ++c) z += exp(p[c] - mx);
                for (int c = 0; c < K; ++c) p[c] = exp(p[c] - mx) / z;
                /* dL/dp_c for dice: -(1/K) * [ 2 t_c (den_c+1) - (2 inter_c + 1) 2 p_c ] / (den_c+1)^2 */
                double dot = 0.0;
                double gp[64];
                if (K > 64) return -1;
                for (int c = 0; c < K; ++c) {
                    const double tc = (c == tg) ? 1.0 : 0.0;
                    const double D1 = den[c] + 1.0;
                    gp[c] = -(2.0 * tc * D1 - (2.0 * inter[c] + 1.0) * 2.0 * p[c]) / (D1 * D1) /
                            (double)K;
                    dot += gp[c] * p[c];
                }
                for (int c = 0; c < K; ++c) {
                    const double tc = (c == tg) ? 1.0 : 0.0;
                    dlogits[((int64_t)b * K + c) * HW + px] =
                        w_dice * p[c] * (gp[c] - dot) + w_ce * (p[c] - tc) / (double)nvalid;
                }
            }
    }
    free(inter); free(den); free(p);
    return 0;
}
