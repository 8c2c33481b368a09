/*
 * sg3_b200.h -- C ABI of libsg3_b200.so: the B200 (sm_100a) kernels behind the
 * StyleGAN3 synthesis hot path of krylea/stylegan3-editing.
 *
 * Every entry point takes plain device/host pointers, sizes and a cudaStream_t
 * passed as void*; no torch types.  All functions are re-entrant and keep no
 * device-global state (the reference's g_fbuf/c_fbuf filter staging,
 * torch_utils/ops/filtered_lrelu.cu:74-84, is replaced by taps passed by value).
 *
 * Return convention (mirrors return_code of filtered_lrelu.cpp:52-56):
 *    0  success, kernel(s) enqueued on `stream`
 *   <0  SG3_E_*: bad arguments, or "no specialised kernel for these parameters"
 *       (SG3_E_NOKERNEL; the caller composes upfirdn2d + act + upfirdn2d like
 *       filtered_lrelu.py:224-230)
 *   >0  a cudaError_t raised while launching
 *
 * Each declaration names the reference pybind entry point it replaces.
 */
#ifndef SG3_B200_H
#define SG3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG3_ABI_VERSION 1

/* element types of activations */
#define SG3_F32 0
#define SG3_F16 1
#define SG3_F64 2   /* bias_act / upfirdn2d / filtered_lrelu_act only */

#define SG3_E_INVALID   (-1)  /* malformed arguments */
#define SG3_E_NOKERNEL  (-2)  /* valid, but no fused specialisation: use the generic composition */
#define SG3_E_TOOLARGE  (-3)  /* exceeds an indexing limit of this entry point */

/* sign-tensor modes of the fused op (filtered_lrelu.cpp:146-160: signWrite / signRead / none) */
#define SG3_SIGNS_NONE  0
#define SG3_SIGNS_WRITE 1
#define SG3_SIGNS_READ  2

int         sg3_abi_version(void);
const char* sg3_error_string(int code);       /* static string for SG3_E_* / cudaError_t codes */
const char* sg3_build_info(void);             /* "sm_100a nvcc x.y ..." */
unsigned long long sg3_launch_count(void);    /* kernels launched by this library since load (all threads) */

/* ------------------------------------------------------------------------
 * filtered_lrelu: bias -> zero-insert upsample + FIR -> gain*lrelu, clamp (+2-bit
 * sign codes) -> FIR + decimate, one fused kernel.
 * Replaces: filtered_lrelu_plugin.filtered_lrelu   (filtered_lrelu.cpp:16-209, .cu:139-1099)
 *
 * Tensors are [N][C][H][W] with BYTE strides (like filtered_lrelu.cpp:127-130); x and y
 * have the same dtype; b is [C] of that dtype (stride in bytes) or NULL.
 * fu / fd are HOST pointers to float32 taps: fuH == 0 means separable with fuW taps
 * (applied along x then y), else a dense [fuH][fuW] filter; same for fd.  NULL = 1x1 identity.
 * Taps are copied into the launch parameters, so the arrays may be freed on return.
 * signs: uint8 [N][C][sH][sWb], 4 pixels per byte along x, pixel k at bits 2k..2k+1,
 * code 1 = negative (scaled by slope), 2 = clamped (gradient 0)   (.cu:494-519).
 * ---------------------------------------------------------------------- */
#define SG3_FLRELU_ROUND_TF32 1

typedef struct sg3_flrelu_desc {
    const void* x;   void* y;   const void* b;   uint8_t* signs;
    const float* fu; const float* fd;            /* host */
    int32_t N, C, inH, inW, outH, outW;
    int64_t xStride[4];                          /* bytes: n, c, h, w */
    int64_t yStride[4];
    int64_t bStride;                             /* bytes */
    int32_t up, down;
    int32_t fuW, fuH, fdW, fdH;                  /* fuH/fdH == 0: separable */
    int32_t px0, py0;                            /* left / top padding in the upsampled domain */
    float   gain, slope, clamp;                  /* clamp = +inf disables */
    int32_t flip;                                /* 1 = correlation, 0 = true convolution */
    int32_t signMode;                            /* SG3_SIGNS_* */
    int32_t sH, sWb;                             /* sign tensor height, width in BYTES */
    int32_t sx, sy;                              /* sign offset added to upsampled coords */
    int32_t dtype;                               /* SG3_F32 / SG3_F16 */
    int32_t flags;                               /* bit 0 (SG3_FLRELU_ROUND_TF32, f32 only): round every output to the nearest TF32 value
                                                    (ties away, as cvt.rna).  For outputs that feed a TF32 tensor-core convolution: the
                                                    tensor core TRUNCATES fp32 operands to TF32, rounding in the producer halves that
                                                    error (what cuDNN's TF32 kernels do on load).  Honoured by the fused forward kernels. */
    float*  ysum;                                /* optional f32 [C]: the kernel ADDS sum_{n,h,w} y[n][c][h][w] (fp32 atomics;
                                                    zero it first).  In the backward pass y = dx and this is the bias gradient
                                                    (filtered_lrelu.py:268 computes dx.sum([0,2,3]) in a second pass).  NULL = off;
                                                    ignored by the pointwise (1x1 filter) kernel: returns SG3_E_NOKERNEL there. */
} sg3_flrelu_desc;

/* Output and sign-tensor geometry (filtered_lrelu.cpp:69-93).  Any out pointer may be NULL. */
int sg3_filtered_lrelu_shape(int inH, int inW, int up, int down,
                             int fuW, int fuH, int fdW, int fdH,
                             int px0, int px1, int py0, int py1,
                             int* outH, int* outW, int* sH, int* sWb);

/* 0 if sg3_filtered_lrelu has a fused kernel for (up, down, filter shapes), SG3_E_NOKERNEL otherwise. */
int sg3_filtered_lrelu_supported(int up, int down, int fuW, int fuH, int fdW, int fdH);

int sg3_filtered_lrelu(const sg3_flrelu_desc* d, void* stream);

/* sizeof(sg3_flrelu_desc) as compiled into the library, so FFI hosts can verify their struct layout. */
int sg3_sizeof_flrelu_desc(void);

/* In-place x = clamp(lrelu(x*gain)) with sign write / read at offset (sx, sy).
 * Replaces: filtered_lrelu_plugin.filtered_lrelu_act_   (filtered_lrelu.cpp:213-290, .cu:1105-1211)
 * x strides in ELEMENTS (n, c, h, w); signs contiguous [N][C][sH][sWb]. */
int sg3_filtered_lrelu_act(void* x, uint8_t* signs,
                           int N, int C, int H, int W, const int64_t xStride[4],
                           int sH, int sWb, int sx, int sy,
                           float gain, float slope, float clamp,
                           int signMode, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * bias_act: y = clamp(act(x + b) * gain) and its first / second derivatives.
 * Replaces: bias_act_plugin.bias_act   (bias_act.cpp:32-90, bias_act.cu:23-147)
 * x, xref, yref, dy, y: dense, same layout, sizeX elements; b: [sizeB] or NULL,
 * channel of element i is (i / stepB) % sizeB.  grad in {0,1,2}; act in 1..9
 * (linear, relu, lrelu, tanh, sigmoid, elu, selu, softplus, swish); clamp < 0 disables.
 * ---------------------------------------------------------------------- */
int sg3_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                 int64_t sizeX, int32_t sizeB, int64_t stepB,
                 int grad, int act, float alpha, float gain, float clamp,
                 int dtype, void* stream);

/* ------------------------------------------------------------------------
 * upfirdn2d: pad / zero-insert upsample / 2-D FIR / decimate.
 * Replaces: upfirdn2d_plugin.upfirdn2d   (upfirdn2d.cpp:16-98, upfirdn2d.cu:29-375)
 * x [N][C][inH][inW], y [N][C][outH][outW], strides in ELEMENTS (n, c, h, w);
 * f: HOST pointer to a dense [fH][fW] float32 filter (a separable filter is two calls).
 * outW = (inW*upx + padx0 + padx1 - fW + downx) / downx      (upfirdn2d.cpp:35-36)
 * ---------------------------------------------------------------------- */
int sg3_upfirdn2d(const void* x, void* y, const float* f,
                  int N, int C, int inH, int inW, int outH, int outW,
                  const int64_t xStride[4], const int64_t yStride[4],
                  int fW, int fH, int upx, int upy, int downx, int downy,
                  int padx0, int pady0, int flip, float gain,
                  int dtype, void* stream);

/* Separable filter with the same up / down factor on both axes in ONE pass over HBM (the reference and sg3_upfirdn2d run
 * it as two launches with an intermediate image, upfirdn2d.py:241-246).  fx [fW], fy [fH]: host taps (NULL = 1);
 * power-of-two factor 1 / 2 / 4 on one side, <= 24 taps per polyphase branch, W-contiguous f32 / f16 tensors; otherwise
 * SG3_E_NOKERNEL (call sg3_upfirdn2d twice).  f32 tensors with factor 1 or 2, <= 12 taps per branch and rows aligned to 8 bytes run the
 * warp-streaming kernel (upfirdn2d_stream.cu: every byte of x and y crosses HBM once, 0.68-0.94 of the copy bandwidth); the rest a CTA-tile kernel. */
int sg3_upfirdn2d_sep(const void* x, void* y, const float* fx, const float* fy,
                      int N, int C, int inH, int inW, int outH, int outW,
                      const int64_t xStride[4], const int64_t yStride[4],
                      int fW, int fH, int up, int down, int padx0, int pady0, int flip, float gain,
                      int dtype, void* stream);

/* ------------------------------------------------------------------------
 * modulated_conv2d (networks_stylegan3.py:24-63).  The reference has no native entry
 * point here: it runs ~10 eager elementwise kernels and a cuDNN grouped convolution.
 *
 * sg3_modconv_weights: fused weight prologue (:39-56) ->
 *   wmod[n][o][i][kh][kw] = w*rsqrt(mean w^2) * s*rsqrt(mean s^2) * rsqrt(sum(.)^2+1e-8) * input_gain
 * w [O][I][k][k] f32, s [N][I] f32, input_gain: NULL, or f32 with gainMode 1 = scalar,
 * 2 = [I], 3 = [N][I].  wmod is f32 [N][O][ldw] with row pitch ldw >= I*k*k floats (zero padded;
 * the tensor-core path needs ldw % 4 == 0); round_tf32 = 1: each value is rounded to the nearest TF32 so the
 * tensor-core contraction sees unbiased operands; round_tf32 = 2: wmod is written as float16 [N][O][ldw] (ldw % 8 == 0),
 * the weight operand of the fp16 tensor-core contraction (what the reference's w.to(x.dtype) produces for fp16 layers); with
 * transpose = 2 as float16 [N][k*k][O][ldw >= I], the operand of the fp16 3x3 kernel;
 * round_tf32 = 3 (layout 0): wmod is f32 [N][2][O][ldw], plane 0 = the weights rounded to TF32, plane 1 = the TF32-rounded
 * residual -- the weight operand of the 3xTF32 contraction (mathMode 2);  round_tf32 = 4: like 1, with the weights first scaled by
 * 1 + 3.52e-4 = the expected relative loss of an fp32 activation that the tensor core truncates to TF32 (2^-11 / (2 ln 2) for a
 * log-uniform mantissa): removes the systematic part of the TF32 operand error for free (any layout that 1 accepts).
 * scratch: >= 4 bytes of device memory (batch-global style norm).
 *
 * sg3_modconv_fwd: y[n][o][p] = sum_{i,tap} wmod[n][o][i][tap] * x[n][i][p + tap - pad]
 * x [N][I][H][W] contiguous, y [N][O][H+2pad-k+1][W+2pad-k+1] contiguous, dtype f32.
 * mathMode 0: FP32 SIMT (exact fp32 accumulate);  1: TF32 tcgen05 implicit GEMM;  2: 3xTF32 tcgen05 GEMM (k = 1 only: both
 * operands split into TF32 head + tail, three MMAs per K step, fp32-accurate to ~1e-6; wmod from round_tf32 = 3).
 * dtype f16 (mathMode 1; k = 1 with H*W % 8 == 0, or k = 3 with pad 0 / 2 and a row pitch of x that is a multiple of 8 -- pass it
 * through sg3_modconv_fwd_pitched when W is not): x, y and wmod are float16, kind::f16 MMAs with fp32 accumulation.
 * ---------------------------------------------------------------------- */
int sg3_modconv_weights(const float* w, const float* s, const float* input_gain, int gainMode,
                        float* wmod, float* scratch,
                        int N, int I, int O, int k, int ldw, int demodulate, int round_tf32, int transpose, void* stream);

/* Weight layouts (`transpose` argument): 0 = [N][O][ldw >= I*k*k] (index i*k*k + tap), read by mathMode 0 and by the
 * 1x1 tensor-core kernel;  1 (1x1 kernels) = transposed [N][I][ldw >= O], the weight operand of the input-gradient GEMM
 * dX = sg3_modconv_fwd(dY, wmodT, ...) with the roles of I and O swapped;  2 = tap-major [N][k*k][O][ldw >= I]
 * (tap = ky*k + kx, i contiguous), the operand sg3_modconv_fwd reads when mathMode = 1 and k > 1;  3 = the same with flipped taps
 * and transposed channels, [N][k*k][I][ldw >= O] with wmod[n][k*k-1-tap][i][o] = W[n][o][i][tap] (buffer zeroed by the caller):
 * the weight operand of the k x k input-gradient conv (roles of I and O swapped).
 *
 * sg3_modconv_tc_supported: 0 if sg3_modconv_fwd(mathMode = 1) has a tensor-core kernel for this shape
 * (k = 1: pad 0, H*W % 4 == 0;  k = 3: pad 0 or 2, W % 4 == 0), SG3_E_NOKERNEL otherwise -- ask before choosing the
 * weight layout.
 *
 * sg3_modconv_wgrad (1x1 kernels, TF32 tcgen05, split-K with fp32 atomics): dw[n][o][i] += sum_p dy[n][o][p] * x[n][i][p];
 * dw [N][O][ldw] must be zeroed by the caller. */
int sg3_modconv_wgrad(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int ldw, void* stream);

/* The same for 3x3 kernels (pad 0 or 2; TF32 tcgen05, the row range split across CTAs, fp32 atomics) -- replaces the grouped
 * weight-gradient convolution the reference runs through conv2d_gradfix.py:153-174 for networks_stylegan3.py:59-62:
 *   dw[n][ky*3+kx][o][i] += sum_oy sum_ox dy[n][o][oy][ox] * x[n][i][oy+ky-pad][ox+kx-pad]      (x = 0 outside the image)
 * dy [N][O][OH][dyPitch], x [N][I][H][xPitch] with OH = H + 2 pad - 2; row pitches in floats, multiples of 4, 0 = dense (then
 * the width itself must be a multiple of 4); dw [N][9][O][ldw >= I] (tap-major like weight layout 2; ldw % 4 == 0, 16-byte aligned)
 * zeroed by the caller.
 * SG3_E_NOKERNEL when a pitch or base pointer is not TMA-addressable. */
int sg3_modconv_wgrad3(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int pad, int ldw,
                       int dyPitch, int xPitch, void* stream);

int sg3_modconv_tc_supported(int I, int O, int H, int W, int k, int pad);

/* Backward of sg3_modconv_weights for 1x1 kernels (the chain rule of networks_stylegan3.py:39-56, ~35 eager kernels per layer
 * when left to autograd): dwmod [N][O][ldw] = gradient wrt the per-sample weights (e.g. from sg3_modconv_wgrad) ->
 * dw [O][I], ds [N][I].  scratch: >= (1 + N*I) floats of device memory.  I <= 2048, otherwise SG3_E_NOKERNEL. */
int sg3_modconv_weights_bwd(const float* dwmod, const float* w, const float* s, const float* input_gain, int gainMode,
                            float* dw, float* ds, float* scratch, int N, int I, int O, int ldw, int demodulate, void* stream);

/* The same chain rule for k x k kernels (3x3: config T), reading the TAP-MAJOR gradient sg3_modconv_wgrad3 accumulates:
 * dwmod [N][k*k][O][ldw >= I] -> dw [O][I][k][k], ds [N][I].  scratch: >= (1 + N*I) floats.  I*k*k <= 4608 (512 channels x 9 taps),
 * otherwise SG3_E_NOKERNEL (the caller then differentiates the weight expression with autograd, as the reference does). */
int sg3_modconv_weights_bwd_taps(const float* dwmod, const float* w, const float* s, const float* input_gain, int gainMode,
                                 float* dw, float* ds, float* scratch, int N, int I, int O, int k, int ldw, int demodulate, void* stream);

int sg3_modconv_fwd(const void* x, const float* wmod, void* y,
                    int N, int I, int O, int H, int W, int k, int pad, int ldw,
                    int mathMode, int dtype, void* stream);

/* Same contraction with padded row pitches: x is [N][I][H][xPitch >= W], y is [N][O][OH][yPitch >= OW] (0 = contiguous).
 * yPitch: first step of the conv -> filtered_lrelu fusion: with a 16-byte-multiple pitch the stencil that
 * consumes y stages its input by TMA even when OW * 4 is not a multiple of 16 (every 3x3 layer of config T: OW = 38 ... 1046).
 * xPitch: lets the same kernel compute the 3x3 INPUT GRADIENT from a dy whose width is not a multiple of 4 (dy comes out of the
 * filtered_lrelu backward kernel with a padded pitch): dX = sg3_modconv_fwd_pitched(dY, wmod in layout 3, pad' = 2 - pad).
 * Implemented by the 3x3 tensor-core kernels (mathMode 1, k = 3; f32 pitches are multiples of 4 elements, f16 pitches of 8);
 * other paths answer SG3_E_NOKERNEL for real pitches. */
int sg3_modconv_fwd_pitched(const void* x, const float* wmod, void* y,
                            int N, int I, int O, int H, int W, int k, int pad, int ldw, int xPitch, int yPitch,
                            int mathMode, int dtype, void* stream);

/* Co-scheduling knob (no reference counterpart; networks.PipelinedSynthesis uses it).  The tensor-core conv kernels are persistent,
 * one CTA per SM with a deep TMA ring (up to ~200 KB of dynamic shared memory), so nothing else fits on an SM while they run.  With a
 * budget of `bytes` (> 0) they shrink the ring to fit, which leaves room for three of the four resident filtered_lrelu CTAs of an SM:
 * launched on two streams, the HBM/tensor-bound contraction of one micro-batch then overlaps the FP32-pipe-bound stencil of another.
 * 0 restores the default.  Process-wide; returns the previous value. */
int sg3_modconv_set_smem_budget(int bytes);

#ifdef __cplusplus
}
#endif
#endif /* SG3_B200_H */
