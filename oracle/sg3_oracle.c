/*
 * sg3_oracle.c -- CPU oracle for the StyleGAN3 synthesis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product path
 * (stylegan3-editing_b200/) never links or imports anything under oracle/.
 *
 * It restates, in plain C with OpenMP over (sample, channel) planes, the
 * algorithm of the reference's impl='ref' ops:
 *
 *   orc_upfirdn2d      <- torch_utils/ops/upfirdn2d.py:168-212   (_upfirdn2d_ref)
 *   orc_bias_act       <- torch_utils/ops/bias_act.py:22-32,92-121 (_bias_act_ref)
 *   orc_lrelu_act      <- torch_utils/ops/filtered_lrelu.cu:1105-1211
 *                         (gain/lrelu/clamp + 2-bit sign codes; the same codes
 *                          as filtered_lrelu.cu:494-505)
 *   orc_modconv2d      <- models/stylegan3/networks_stylegan3.py:24-63
 *                         (modulated_conv2d: pre-norm, modulate, demodulate,
 *                          input gain, grouped conv with zero padding)
 *   orc_conv2d_*       <- the F.conv2d call at networks_stylegan3.py:61 and its
 *                         autograd transposes (dgrad / wgrad)
 *
 * The filtered_lrelu composition itself (filtered_lrelu.py:122-154) and its
 * backward (filtered_lrelu.py:240-269) are assembled from these primitives in
 * oracle/sg3_oracle.py.
 *
 * Numerics: tensors are float32 in memory like the reference's fp32 CPU path;
 * every reduction accumulates in double and rounds once to float32 at the op
 * boundary, which makes the result independent of thread count and summation
 * order.  Parity pin: tests/test_oracle_golden.py checks this file against
 * outputs of the real reference (imported from /root/reference by
 * tests/golden/make_golden.py) -- the reference ships no tests or golden
 * vectors of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* ------------------------------------------------------------------------
 * upfirdn2d, one dense 2-D filter f[fH][fW] (a separable filter is two calls
 * with [1][n] and [n][1] filters, exactly as upfirdn2d.py:205-208 runs two
 * depthwise convolutions).
 *
 *   Z[pady0 + upy*i][padx0 + upx*j] = x[i][j], zero elsewhere    (:188-194)
 *   F' = f flipped in both axes unless flip_filter               (:199-200)
 *   y[oy][ox] = gain * sum_{a,b} F'[a][b] * Z[oy*downy + a][ox*downx + b]   (:203-211)
 *
 * gain is applied to the taps (f * gain) before the sum like :197.
 * ---------------------------------------------------------------------- */
ORC_API int orc_upfirdn2d(const float* x, float* y, int64_t planes, int inH, int inW,
                          const float* f, int fH, int fW,
                          int upx, int upy, int downx, int downy,
                          int padx0, int padx1, int pady0, int pady1,
                          int flip_filter, double gain)
{
    const int ZW = inW * upx + padx0 + padx1;
    const int ZH = inH * upy + pady0 + pady1;
    if (ZW < fW || ZH < fH) return -1;
    const int outW = (ZW - fW + downx) / downx;
    const int outH = (ZH - fH + downy) / downy;

    /* Effective correlation taps, rounded to float32 like `f * gain ** (ndim/2)` then `.to(x.dtype)`. */
    float* taps = (float*)malloc(sizeof(float) * (size_t)fH * fW);
    for (int a = 0; a < fH; a++)
        for (int b = 0; b < fW; b++) {
            int sa = flip_filter ? a : fH - 1 - a;
            int sb = flip_filter ? b : fW - 1 - b;
            taps[a * fW + b] = (float)((double)f[sa * fW + sb] * gain);
        }

#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t p = 0; p < planes; p++) {
        const float* xp = x + p * (int64_t)inH * inW;
        float* yp = y + p * (int64_t)outH * outW;
        for (int oy = 0; oy < outH; oy++) {
            /* first tap row a0 >= 0 with (oy*downy + a0 - pady0) % upy == 0 */
            int zy0 = oy * downy - pady0;
            int a0 = ((-zy0) % upy + upy) % upy;
            for (int ox = 0; ox < outW; ox++) {
                int zx0 = ox * downx - padx0;
                int b0 = ((-zx0) % upx + upx) % upx;
                double acc = 0.0;
                for (int a = a0; a < fH; a += upy) {
                    int i = (zy0 + a) / upy;          /* exact: zy0 + a is a multiple of upy */
                    if (zy0 + a < 0 || i >= inH) continue;
                    const float* xr = xp + (int64_t)i * inW;
                    const float* tr = taps + a * fW;
                    for (int b = b0; b < fW; b += upx) {
                        int j = (zx0 + b) / upx;
                        if (zx0 + b < 0 || j >= inW) continue;
                        acc += (double)tr[b] * (double)xr[j];
                    }
                }
                yp[(int64_t)oy * outW + ox] = (float)acc;
            }
        }
    }
    free(taps);
    return 0;
}

/* ------------------------------------------------------------------------
 * bias_act forward: y = clamp(act(x + b[c]) * gain)     bias_act.py:92-121
 * act index follows the table at bias_act.py:22-32 (1=linear ... 9=swish).
 * x viewed as [outer][C][inner]; b may be NULL; clamp < 0 disables clamping.
 * ---------------------------------------------------------------------- */
static inline double act_fwd(int act, double v, double alpha)
{
    switch (act) {
    case 1: return v;
    case 2: return v > 0 ? v : 0.0;
    case 3: return v > 0 ? v : v * alpha;
    case 4: return tanh(v);
    case 5: return 1.0 / (1.0 + exp(-v));
    case 6: return v > 0 ? v : expm1(v);
    case 7: return 1.0507009873554804934193349852946 * (v > 0 ? v : 1.6732632423543772848170429916717 * expm1(v));
    case 8: return v > 20.0 ? v : log1p(exp(v));          /* F.softplus threshold=20 */
    case 9: return v / (1.0 + exp(-v));
    default: return NAN;
    }
}

ORC_API int orc_bias_act(const float* x, const float* b, float* y,
                         int64_t outer, int64_t C, int64_t inner,
                         int act, double alpha, double gain, double clamp)
{
    if (act < 1 || act > 9) return -1;
    const int64_t total = outer * C * inner;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < total; idx++) {
        int64_t c = (idx / inner) % C;
        float xb = b ? (float)(x[idx] + b[c]) : x[idx];          /* x + b rounded to fp32 first (:106) */
        float v = (float)act_fwd(act, (double)xb, alpha);          /* activation output is an fp32 tensor (:110) */
        if (gain != 1.0) v = (float)((double)v * gain);            /* :113-115 */
        if (clamp >= 0.0) {                                        /* :118-119 */
            if ((double)v > clamp) v = (float)clamp;
            if ((double)v < -clamp) v = (float)-clamp;
        }
        y[idx] = v;
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * In-place gain / leaky-ReLU / clamp with the 2-bit sign code of the fused op.
 *   mode 0: no signs          v = clamp(lrelu(v*gain))
 *   mode 1: write signs       same, and code = 1 if v*gain < 0, 2 if |lrelu| > clamp
 *   mode 2: read signs        v = v*gain * {1, slope, 0}[code at (X+sx, Y+sy)], untouched
 *                             outside the sign window
 * x is [planes][H][W]; signs is [planes][sH][sWb] bytes, 4 pixels per byte along
 * x, pixel k of a byte at bits 2k..2k+1          (filtered_lrelu.cu:494-519,1127-1190)
 * ---------------------------------------------------------------------- */
ORC_API int orc_lrelu_act(float* x, uint8_t* signs, int64_t planes, int H, int W,
                          int sH, int sWb, int sx, int sy,
                          double gain, double slope, double clamp, int mode)
{
    if (mode < 0 || mode > 2) return -1;
    if (mode == 1) memset(signs, 0, (size_t)planes * sH * sWb);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < planes; p++) {
        float* xp = x + p * (int64_t)H * W;
        uint8_t* sp = signs ? signs + p * (int64_t)sH * sWb : NULL;
        for (int yy = 0; yy < H; yy++)
            for (int xx = 0; xx < W; xx++) {
                float v = (float)((double)xp[(int64_t)yy * W + xx] * (float)gain);
                if (mode == 2) {
                    int px = xx + sx, py = yy + sy;
                    if (px >= 0 && py >= 0 && py < sH && (px >> 2) < sWb) {
                        int code = (sp[(int64_t)py * sWb + (px >> 2)] >> ((px & 3) * 2)) & 3;
                        if (code & 1) v = (float)((double)v * (float)slope);
                        if (code & 2) v = 0.0f;
                    }
                } else {
                    int code = 0;
                    if (v < 0.0f) { v = (float)((double)v * (float)slope); code = 1; }
                    if (fabs((double)v) > clamp) { v = (float)(v < 0 ? -clamp : clamp); code = 2; }
                    if (mode == 1) {
                        int px = xx + sx, py = yy + sy;
                        if (px >= 0 && py >= 0 && py < sH && (px >> 2) < sWb)
                            sp[(int64_t)py * sWb + (px >> 2)] |= (uint8_t)(code << ((px & 3) * 2));
                    }
                }
                xp[(int64_t)yy * W + xx] = v;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * Plain cross-correlation with zero padding, per-sample weights (groups = N):
 *   y[n][o][oy][ox] = sum_{i,a,b} w[n][o][i][a][b] * x[n][i][oy + a - pad][ox + b - pad]
 * This is F.conv2d(groups=batch) of networks_stylegan3.py:59-62.
 * If shared_w != 0, w is [O][I][k][k] and used for every sample.
 * ---------------------------------------------------------------------- */
ORC_API int orc_conv2d_fwd(const float* x, const float* w, float* y,
                           int N, int I, int O, int H, int W, int k, int pad, int shared_w)
{
    const int OH = H + 2 * pad - k + 1, OW = W + 2 * pad - k + 1;
    if (OH <= 0 || OW <= 0) return -1;
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)OH * OW);
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int n = 0; n < N; n++)
            for (int o = 0; o < O; o++) {
                memset(acc, 0, sizeof(double) * (size_t)OH * OW);
                const float* wn = w + (shared_w ? 0 : (int64_t)n * O * I * k * k) + (int64_t)o * I * k * k;
                for (int i = 0; i < I; i++) {
                    const float* xi = x + ((int64_t)n * I + i) * H * W;
                    for (int a = 0; a < k; a++)
                        for (int b = 0; b < k; b++) {
                            const double wv = (double)wn[(i * k + a) * k + b];
                            int oy0 = pad - a > 0 ? pad - a : 0;
                            int oy1 = H + pad - a < OH ? H + pad - a : OH;
                            int ox0 = pad - b > 0 ? pad - b : 0;
                            int ox1 = W + pad - b < OW ? W + pad - b : OW;
                            for (int oy = oy0; oy < oy1; oy++) {
                                const float* xr = xi + (int64_t)(oy + a - pad) * W + (b - pad);
                                double* ar = acc + (int64_t)oy * OW;
                                for (int ox = ox0; ox < ox1; ox++) ar[ox] += wv * (double)xr[ox];
                            }
                        }
                }
                float* yo = y + ((int64_t)n * O + o) * OH * OW;
                for (int64_t q = 0; q < (int64_t)OH * OW; q++) yo[q] = (float)acc[q];
            }
        free(acc);
    }
    return 0;
}

/* dgrad: dx[n][i][iy][ix] = sum_{o,a,b} w[n][o][i][a][b] * dy[n][o][iy - a + pad][ix - b + pad] */
ORC_API int orc_conv2d_dgrad(const float* dy, const float* w, float* dx,
                             int N, int I, int O, int H, int W, int k, int pad, int shared_w)
{
    const int OH = H + 2 * pad - k + 1, OW = W + 2 * pad - k + 1;
#pragma omp parallel
    {
        double* acc = (double*)malloc(sizeof(double) * (size_t)H * W);
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int n = 0; n < N; n++)
            for (int i = 0; i < I; i++) {
                memset(acc, 0, sizeof(double) * (size_t)H * W);
                for (int o = 0; o < O; o++) {
                    const float* wn = w + (shared_w ? 0 : (int64_t)n * O * I * k * k) + ((int64_t)o * I + i) * k * k;
                    const float* dyo = dy + ((int64_t)n * O + o) * OH * OW;
                    for (int a = 0; a < k; a++)
                        for (int b = 0; b < k; b++) {
                            const double wv = (double)wn[a * k + b];
                            for (int iy = 0; iy < H; iy++) {
                                int oy = iy - a + pad;
                                if (oy < 0 || oy >= OH) continue;
                                for (int ix = 0; ix < W; ix++) {
                                    int ox = ix - b + pad;
                                    if (ox < 0 || ox >= OW) continue;
                                    acc[(int64_t)iy * W + ix] += wv * (double)dyo[(int64_t)oy * OW + ox];
                                }
                            }
                        }
                }
                float* dxi = dx + ((int64_t)n * I + i) * H * W;
                for (int64_t q = 0; q < (int64_t)H * W; q++) dxi[q] = (float)acc[q];
            }
        free(acc);
    }
    return 0;
}

/* wgrad (per sample): dw[n][o][i][a][b] = sum_{oy,ox} dy[n][o][oy][ox] * x[n][i][oy + a - pad][ox + b - pad] */
ORC_API int orc_conv2d_wgrad(const float* x, const float* dy, float* dw,
                             int N, int I, int O, int H, int W, int k, int pad)
{
    const int OH = H + 2 * pad - k + 1, OW = W + 2 * pad - k + 1;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int n = 0; n < N; n++)
        for (int o = 0; o < O; o++) {
            const float* dyo = dy + ((int64_t)n * O + o) * OH * OW;
            for (int i = 0; i < I; i++) {
                const float* xi = x + ((int64_t)n * I + i) * H * W;
                for (int a = 0; a < k; a++)
                    for (int b = 0; b < k; b++) {
                        double acc = 0.0;
                        for (int oy = 0; oy < OH; oy++) {
                            int iy = oy + a - pad;
                            if (iy < 0 || iy >= H) continue;
                            for (int ox = 0; ox < OW; ox++) {
                                int ix = ox + b - pad;
                                if (ix < 0 || ix >= W) continue;
                                acc += (double)dyo[(int64_t)oy * OW + ox] * (double)xi[(int64_t)iy * W + ix];
                            }
                        }
                        dw[((((int64_t)n * O + o) * I + i) * k + a) * k + b] = (float)acc;
                    }
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------
 * modulated_conv2d weights                         networks_stylegan3.py:39-56
 *   w' = w * rsqrt(mean_{i,a,b} w^2)        (per o)       if demodulate  (:41)
 *   s' = s * rsqrt(mean_{n,i} s^2)          (batch-global) if demodulate (:42)
 *   W[n,o,i,a,b] = w'[o,i,a,b] * s'[n,i]                                  (:45-46)
 *   d[n,o] = rsqrt(sum_{i,a,b} W^2 + 1e-8);  W *= d        if demodulate  (:49-51)
 *   W *= input_gain[n,i]   (gain_mode 0 none / 1 scalar / 2 [I] / 3 [N][I]) (:54-56)
 * Each step rounds to float32 like the chain of fp32 tensor ops.
 * wmod is [N][O][I][k][k].
 * ---------------------------------------------------------------------- */
ORC_API int orc_modconv_weights(const float* w, const float* s, const float* input_gain, int gain_mode,
                                float* wmod, int N, int I, int O, int k, int demodulate)
{
    const int kk = k * k;
    float* wn = (float*)malloc(sizeof(float) * (size_t)O * I * kk);
    float* sn = (float*)malloc(sizeof(float) * (size_t)N * I);
    if (demodulate) {
        for (int o = 0; o < O; o++) {
            double acc = 0.0;
            for (int q = 0; q < I * kk; q++) { float v = w[(int64_t)o * I * kk + q]; acc += (double)(float)(v * v); }
            float r = (float)(1.0 / sqrt((double)(float)(acc / (double)(I * kk))));
            for (int q = 0; q < I * kk; q++) wn[(int64_t)o * I * kk + q] = w[(int64_t)o * I * kk + q] * r;
        }
        double acc = 0.0;
        for (int q = 0; q < N * I; q++) acc += (double)(float)(s[q] * s[q]);
        float r = (float)(1.0 / sqrt((double)(float)(acc / (double)(N * I))));
        for (int q = 0; q < N * I; q++) sn[q] = s[q] * r;
    } else {
        memcpy(wn, w, sizeof(float) * (size_t)O * I * kk);
        memcpy(sn, s, sizeof(float) * (size_t)N * I);
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int o = 0; o < O; o++) {
            float* dst = wmod + ((int64_t)n * O + o) * I * kk;
            double acc = 0.0;
            for (int i = 0; i < I; i++)
                for (int q = 0; q < kk; q++) {
                    float v = wn[((int64_t)o * I + i) * kk + q] * sn[n * I + i];
                    dst[i * kk + q] = v;
                    acc += (double)(float)(v * v);
                }
            if (demodulate) {
                float d = (float)(1.0 / sqrt((double)((float)acc + 1e-8f)));
                for (int q = 0; q < I * kk; q++) dst[q] = dst[q] * d;
            }
            if (gain_mode) {
                for (int i = 0; i < I; i++) {
                    float g = gain_mode == 1 ? input_gain[0] : gain_mode == 2 ? input_gain[i] : input_gain[n * I + i];
                    for (int q = 0; q < kk; q++) dst[i * kk + q] = dst[i * kk + q] * g;
                }
            }
        }
    free(wn);
    free(sn);
    return 0;
}
