// modconv.cu -- modulated_conv2d of the StyleGAN3 synthesis layer
// (models/stylegan3/networks_stylegan3.py:24-63), split into
//   1. a fused weight prologue: pre-normalise, modulate, demodulate, input gain -> per-sample
//      weights wmod[n][o][i*k*k]   (the ~10 eager elementwise/reduction kernels of :39-56);
//   2. the contraction y[n][o][p] = sum_{i,tap} wmod[n][o][i][tap] * x[n][i][p + tap - pad]
//      (the cuDNN grouped convolution of :59-62).
// This file holds the prologue and the exact-FP32 SIMT contraction (mathMode 0) used for
// fp32 parity; the TF32 tcgen05/TMEM implicit GEMM (mathMode 1) lives in modconv_tc.cu.
#include "common.cuh"

int sg3_modconv_wgrad_tc(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int ldw, cudaStream_t stream);
int sg3_modconv_wgrad3_tc(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int pad, int ldw,
                          int dyPitch, int xPitch, cudaStream_t stream);
int sg3_modconv_fwd_tc3(const float* x, const float* wtap, float* y, int N, int I, int O, int H, int W, int pad, int ldw,
                        int xPitch, int yPitch, cudaStream_t stream);
int sg3_modconv_fwd_tc3_f16(const void* x, const void* wtap, void* y, int N, int I, int O, int H, int W, int pad, int ldw,
                            int xPitch, int yPitch, cudaStream_t stream);
int sg3_modconv_tc3_supported(int I, int O, int H, int W, int k, int pad);
int sg3_modconv_fwd_tc_f16(const void* x, const void* wmod, void* y, int N, int I, int O, int H, int W, int k, int pad, int ldw, cudaStream_t stream);
int sg3_modconv_fwd_tc_x3(const float* x, const float* wmod, float* y, int N, int I, int O, int H, int W, int k, int pad, int ldw,
                          cudaStream_t stream);
int sg3_modconv_fwd_tc(const float* x, const float* wmod, float* y, int N, int I, int O, int H, int W, int k, int pad, int ldw,
                       cudaStream_t stream);

namespace {

__device__ __forceinline__ float block_sum(float v, float* red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = l < nw ? red[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// scratch[0] = rsqrt(mean(s^2)) over the whole batch          (:42)
__global__ void __launch_bounds__(1024) style_norm_kernel(const float* __restrict__ s, int count, float* scratch)
{
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) { float v = s[i]; acc += v * v; }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) scratch[0] = rsqrtf(acc / (float)count);
}

__device__ __forceinline__ float round_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// One CTA per output channel o: rw = rsqrt(mean w[o]^2), then for every sample n the
// demodulation coefficient and the final weights.
__global__ void __launch_bounds__(256) modconv_weights_kernel(
    const float* __restrict__ w, const float* __restrict__ s, const float* __restrict__ gain, int gainMode,
    float* __restrict__ wmod, const float* __restrict__ scratch,
    int N, int I, int O, int kk, int ldw, int demodulate, int roundTf32, int transpose)
{
    __shared__ float red[32];
    const int o = blockIdx.x;
    const int cnt = I * kk;
    const float* wo = w + (size_t)o * cnt;
    float rw = 1.f, rs = 1.f;
    if (demodulate) {
        float acc = 0.f;
        for (int q = threadIdx.x; q < cnt; q += blockDim.x) { float v = wo[q]; acc += v * v; }
        acc = block_sum(acc, red);
        rw = rsqrtf(acc / (float)cnt);
        rs = scratch[0];
    }
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const float* sn = s + (size_t)n * I;
        float d = 1.f;
        if (demodulate) {
            float acc = 0.f;
            for (int q = threadIdx.x; q < cnt; q += blockDim.x) {
                float v = (wo[q] * rw) * (sn[q / kk] * rs);
                acc += v * v;
            }
            acc = block_sum(acc, red);
            d = rsqrtf(acc + 1e-8f);
        }
        // layout 1 (1x1 kernels only): transposed, wmod is [N][I][ldw >= O], the operand of the dgrad GEMM
        // layout 2: tap-major, wmod is [N][kk][O][ldw >= I], the operand of the 3x3 tensor-core kernel
        // layout 3: tap-major with flipped taps and transposed channels, wmod is [N][kk][I][ldw >= O] with
        //           wmod[n][kk-1-tap][i][o] = W[n][o][i][tap] -- the operand of the 3x3 INPUT-GRADIENT conv (same kernel, roles of I / O swapped)
        float* dst = transpose == 1 ? wmod + (size_t)n * I * ldw + o
                   : transpose == 2 ? wmod + ((size_t)n * kk * O + o) * ldw
                   : transpose == 3 ? wmod + (size_t)n * kk * I * ldw + o
                                    : wmod + ((size_t)n * O + o) * ldw;
        const size_t dstStep = transpose == 1 ? (size_t)ldw : 1;
        const size_t tapStep = (size_t)O * ldw;
        // fp16 operand form: layout 0 (1x1 kernel) or layout 2 (tap-major, the fp16 3x3 kernel)
        __half* dstH = reinterpret_cast<__half*>(wmod) + (transpose == 2 ? ((size_t)n * kk * O + o) * ldw : ((size_t)n * O + o) * ldw);
        if (roundTf32 == 3) {
            // 3xTF32 operand form (layout 0 only): [N][2][O][ldw], plane 0 = TF32 head, plane 1 = TF32 tail of the residual
            float* hi = wmod + ((size_t)(2 * n) * O + o) * ldw;
            float* lo = hi + (size_t)O * ldw;
            for (int q = cnt + threadIdx.x; q < ldw; q += blockDim.x) { hi[q] = 0.f; lo[q] = 0.f; }
            for (int q = threadIdx.x; q < cnt; q += blockDim.x) {
                const int i = q / kk;
                float v = (wo[q] * rw) * (sn[i] * rs);
                if (demodulate) v *= d;
                if (gainMode == 1) v *= gain[0];
                else if (gainMode == 2) v *= gain[i];
                else if (gainMode == 3) v *= gain[(size_t)n * I + i];
                const float h = round_tf32(v);
                hi[q] = h;
                lo[q] = round_tf32(v - h);
            }
            continue;
        }
        if (transpose == 0 && roundTf32 != 2)
            for (int q = cnt + threadIdx.x; q < ldw; q += blockDim.x) dst[q] = 0.f;   // row padding (TMA pitch)
        if (roundTf32 == 2 && transpose == 0)
            for (int q = cnt + threadIdx.x; q < ldw; q += blockDim.x) dstH[q] = __float2half_rn(0.f);
        for (int q = threadIdx.x; q < cnt; q += blockDim.x) {
            const int i = q / kk;
            float v = (wo[q] * rw) * (sn[i] * rs);
            if (demodulate) v *= d;
            if (gainMode == 1) v *= gain[0];
            else if (gainMode == 2) v *= gain[i];
            else if (gainMode == 3) v *= gain[(size_t)n * I + i];
            if (roundTf32 == 2) {                                                // what the reference's w.to(x.dtype) does (:61)
                if (transpose == 2) dstH[(size_t)(q - i * kk) * tapStep + i] = __float2half_rn(v);
                else dstH[q] = __float2half_rn(v);
                continue;
            }
            // 4: the tensor core will TRUNCATE the fp32 activations this weight multiplies (10 of 23 mantissa bits survive): a
            // multiplicative bias of -2^-11 * E[1/m] = -2^-11 / (2 ln 2) = -3.52e-4 for a log-uniform mantissa m, the same sign on
            // every term of the sum.  It is folded into the weight before the weight itself is rounded (to nearest).
            if (roundTf32 == 4) v *= 1.0f + 3.5217e-4f;
            v = roundTf32 ? round_tf32(v) : v;
            if (transpose == 2) dst[(size_t)(q - i * kk) * tapStep + i] = v;
            else if (transpose == 3) dst[((size_t)(kk - 1 - (q - i * kk)) * I + i) * ldw] = v;
            else dst[q * dstStep] = v;
        }
    }
}


// ---- backward of the weight prologue (1x1 kernels): dWf = dL/d wmod [N][O][ldw]  ->  dw [O][I], dsn [N][I] ---------------
// Chain rule of networks_stylegan3.py:39-56 in one pass per output channel (the torch version of this chain is ~35 small
// kernels per layer):  wn = w rw, sn = s rs, Wm = wn sn, q = sum_i Wm^2 + 1e-8, wmod = Wm q^-1/2 g
//   dW = dWf g;  dd = sum_i dW Wm;  dWm = dW q^-1/2 - Wm dd q^-3/2;  dwn = sum_n dWm sn;  dsn = sum_o dWm wn (atomics)
//   dw = rw dwn - w rw^3 (sum_i dwn w) / I.     The style side (ds from dsn, a global reduction) is modconv_style_bwd_kernel.
// One CTA per output channel o; thread t owns input channels t, t + 256, ... (up to kMaxIPerThread of them).
constexpr int kMaxIPerThread = 8;      // I <= 2048

__global__ void __launch_bounds__(256) modconv_weights_bwd_kernel(
    const float* __restrict__ dWf, const float* __restrict__ w, const float* __restrict__ s, const float* __restrict__ gain, int gainMode,
    const float* __restrict__ scratch, float* __restrict__ dw, float* __restrict__ dsn,
    int N, int I, int O, int ldw, int demodulate)
{
    __shared__ float red[32];
    const int o = blockIdx.x;
    const float* wo = w + (size_t)o * I;
    float wreg[kMaxIPerThread], dwn[kMaxIPerThread];
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxIPerThread; u++) {
        const int i = threadIdx.x + 256 * u;
        wreg[u] = i < I ? wo[i] : 0.f;
        dwn[u] = 0.f;
        acc += wreg[u] * wreg[u];
    }
    float rw = 1.f, rs = 1.f;
    if (demodulate) {
        acc = block_sum(acc, red);
        rw = rsqrtf(acc / (float)I);
        rs = scratch[0];
    }
    for (int n = 0; n < N; n++) {
        const float* sn = s + (size_t)n * I;
        const float* dr = dWf + ((size_t)n * O + o) * ldw;
        float Wm[kMaxIPerThread], dW[kMaxIPerThread];
        float q = 0.f, dd = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxIPerThread; u++) {
            const int i = threadIdx.x + 256 * u;
            float g = 1.f;
            if (i < I) g = gainMode == 1 ? gain[0] : gainMode == 2 ? gain[i] : gainMode == 3 ? gain[(size_t)n * I + i] : 1.f;
            Wm[u] = i < I ? (wreg[u] * rw) * (sn[i] * rs) : 0.f;
            dW[u] = i < I ? dr[i] * g : 0.f;
            q += Wm[u] * Wm[u];
            dd += dW[u] * Wm[u];
        }
        float d = 1.f, c = 0.f;
        if (demodulate) {
            q = block_sum(q, red) + 1e-8f;
            dd = block_sum(dd, red);
            d = rsqrtf(q);
            c = dd * d / q;                       // dd q^-3/2
        }
#pragma unroll
        for (int u = 0; u < kMaxIPerThread; u++) {
            const int i = threadIdx.x + 256 * u;
            if (i < I) {
                const float dWm = dW[u] * d - Wm[u] * c;
                dwn[u] += dWm * (sn[i] * rs);
                atomicAdd(dsn + (size_t)n * I + i, dWm * (wreg[u] * rw));
            }
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxIPerThread; u++) dot += dwn[u] * wreg[u];
    if (demodulate) dot = block_sum(dot, red);
#pragma unroll
    for (int u = 0; u < kMaxIPerThread; u++) {
        const int i = threadIdx.x + 256 * u;
        if (i < I) dw[(size_t)o * I + i] = demodulate ? rw * dwn[u] - wreg[u] * rw * rw * rw * dot / (float)I : dwn[u];
    }
}

// The same chain rule for k x k kernels (3x3: config T).  One CTA per output channel; the I * kk weights of the channel are
// flattened as q = i * kk + tap (the layout of w itself); thread t owns q = t, t + 256, ... (up to KPT of them).  The incoming
// gradient is TAP-MAJOR, dWf[n][tap][o][i] with row pitch ldw -- what sg3_modconv_wgrad3 accumulates.  Differences from the 1x1
// kernel: the style / gain index is q / kk, the pre-normalisation runs over I * kk values (:41), nothing else.
template <int KPT>
__global__ void __launch_bounds__(256) modconv_weights_bwd_taps_kernel(
    const float* __restrict__ dWf, const float* __restrict__ w, const float* __restrict__ s, const float* __restrict__ gain, int gainMode,
    const float* __restrict__ scratch, float* __restrict__ dw, float* __restrict__ dsn,
    int N, int I, int O, int kk, int ldw, int demodulate)
{
    __shared__ float red[32];
    const int o = blockIdx.x;
    const int cnt = I * kk;
    const float* wo = w + (size_t)o * cnt;
    float wreg[KPT], dwn[KPT];
    int ii[KPT], src[KPT];                    // input channel of q, and its offset inside one sample's tap-major gradient
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < KPT; u++) {
        const int q = threadIdx.x + 256 * u;
        const int i = q / kk;
        ii[u] = i;
        src[u] = ((q - i * kk) * O + o) * ldw + i;
        wreg[u] = q < cnt ? wo[q] : 0.f;
        dwn[u] = 0.f;
        acc += wreg[u] * wreg[u];
    }
    float rw = 1.f, rs = 1.f;
    if (demodulate) {
        acc = block_sum(acc, red);
        rw = rsqrtf(acc / (float)cnt);
        rs = scratch[0];
    }
    for (int n = 0; n < N; n++) {
        const float* sn = s + (size_t)n * I;
        const float* dr = dWf + (size_t)n * kk * O * ldw;
        float Wm[KPT], dW[KPT];
        float q2 = 0.f, dd = 0.f;
#pragma unroll
        for (int u = 0; u < KPT; u++) {
            const bool ok = threadIdx.x + 256 * u < cnt;
            const int i = ii[u];
            float g = 1.f;
            if (ok) g = gainMode == 1 ? gain[0] : gainMode == 2 ? gain[i] : gainMode == 3 ? gain[(size_t)n * I + i] : 1.f;
            Wm[u] = ok ? (wreg[u] * rw) * (sn[i] * rs) : 0.f;
            dW[u] = ok ? dr[src[u]] * g : 0.f;
            q2 += Wm[u] * Wm[u];
            dd += dW[u] * Wm[u];
        }
        float d = 1.f, c = 0.f;
        if (demodulate) {
            q2 = block_sum(q2, red) + 1e-8f;
            dd = block_sum(dd, red);
            d = rsqrtf(q2);
            c = dd * d / q2;                      // dd q^-3/2
        }
#pragma unroll
        for (int u = 0; u < KPT; u++) {
            if (threadIdx.x + 256 * u < cnt) {
                const int i = ii[u];
                const float dWm = dW[u] * d - Wm[u] * c;
                dwn[u] += dWm * (sn[i] * rs);
                atomicAdd(dsn + (size_t)n * I + i, dWm * (wreg[u] * rw));
            }
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < KPT; u++) dot += dwn[u] * wreg[u];
    if (demodulate) dot = block_sum(dot, red);
#pragma unroll
    for (int u = 0; u < KPT; u++) {
        const int q = threadIdx.x + 256 * u;
        if (q < cnt) dw[(size_t)o * cnt + q] = demodulate ? rw * dwn[u] - wreg[u] * rw * rw * rw * dot / (float)cnt : dwn[u];
    }
}

// ds = rs dsn - s rs^3 (sum_{n,i} dsn s) / (N I)   (demodulate), else ds = dsn.  One CTA.
__global__ void __launch_bounds__(1024) modconv_style_bwd_kernel(const float* __restrict__ dsn, const float* __restrict__ s,
                                                                 const float* __restrict__ scratch, float* __restrict__ ds, int count, int demodulate)
{
    __shared__ float red[32];
    if (!demodulate) {
        for (int i = threadIdx.x; i < count; i += blockDim.x) ds[i] = dsn[i];
        return;
    }
    float acc = 0.f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += dsn[i] * s[i];
    acc = block_sum(acc, red);
    const float rs = scratch[0];
    const float k = rs * rs * rs * acc / (float)count;
    for (int i = threadIdx.x; i < count; i += blockDim.x) ds[i] = rs * dsn[i] - s[i] * k;
}

// ---- exact FP32 contraction: 64 output channels x 64 pixels per CTA, K chunks of 16 --------
constexpr int BO = 64, BP = 64, BK = 16;

__global__ void __launch_bounds__(256) modconv_fwd_simt_kernel(
    const float* __restrict__ x, const float* __restrict__ wmod, float* __restrict__ y,
    int N, int I, int O, int H, int W, int k, int pad, int OH, int OW, int ldw)
{
    __shared__ float sA[BK][BO + 4];   // weights  [kchunk][o]
    __shared__ float sB[BK][BP + 4];   // pixels   [kchunk][p]
    const int P = OH * OW;
    const int kk = k * k;
    const int K = I * kk;
    const int tilesP = (P + BP - 1) / BP, tilesO = (O + BO - 1) / BO;
    const int64_t total = (int64_t)N * tilesO * tilesP;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4x4 outputs each
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = (int)(t / ((int64_t)tilesO * tilesP));
        const int rem = (int)(t - (int64_t)n * tilesO * tilesP);
        const int to = rem / tilesP, tp = rem - to * tilesP;
        const int o0 = to * BO, p0 = tp * BP;
        const float* wn = wmod + (size_t)n * O * ldw;
        const float* xn = x + (size_t)n * I * H * W;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < K; k0 += BK) {
            __syncthreads();
            for (int e = threadIdx.x; e < BK * BO; e += 256) {
                const int o = e / BK, kc = e - o * BK;
                float v = 0.f;
                if (o0 + o < O && k0 + kc < K) v = wn[(size_t)(o0 + o) * ldw + k0 + kc];
                sA[kc][o] = v;
            }
            for (int e = threadIdx.x; e < BK * BP; e += 256) {
                const int kc = e / BP, pp = e - kc * BP;
                float v = 0.f;
                const int kidx = k0 + kc, p = p0 + pp;
                if (kidx < K && p < P) {
                    const int i = kidx / kk, tap = kidx - i * kk;
                    const int ta = tap / k, tb = tap - ta * k;
                    const int oy = p / OW, ox = p - oy * OW;
                    const int iy = oy + ta - pad, ix = ox + tb - pad;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = xn[((size_t)i * H + iy) * W + ix];
                }
                sB[kc][pp] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kc = 0; kc < BK; kc++) {
                float a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; q++) { a[q] = sA[kc][ty * 4 + q]; b[q] = sB[kc][tx * 4 + q]; }
#pragma unroll
                for (int qa = 0; qa < 4; qa++)
#pragma unroll
                    for (int qb = 0; qb < 4; qb++) acc[qa][qb] = fmaf(a[qa], b[qb], acc[qa][qb]);
            }
        }
#pragma unroll
        for (int qa = 0; qa < 4; qa++) {
            const int o = o0 + ty * 4 + qa;
            if (o >= O) continue;
#pragma unroll
            for (int qb = 0; qb < 4; qb++) {
                const int p = p0 + tx * 4 + qb;
                if (p < P) y[((size_t)n * O + o) * P + p] = acc[qa][qb];
            }
        }
    }
}

}  // namespace

SG3_EXPORT int sg3_modconv_weights(const float* w, const float* s, const float* input_gain, int gainMode,
                                   float* wmod, float* scratch,
                                   int N, int I, int O, int k, int ldw, int demodulate, int round_tf32_flag, int transpose, void* stream)
{
    if (!w || !s || !wmod || !scratch || N < 1 || I < 1 || O < 1 || k < 1) return SG3_E_INVALID;
    if (transpose < 0 || transpose > 3 || round_tf32_flag < 0 || round_tf32_flag > 4) return SG3_E_INVALID;
    if (round_tf32_flag == 3 && transpose != 0) return SG3_E_INVALID;
    if (round_tf32_flag == 2 && transpose != 0 && transpose != 2) return SG3_E_INVALID;
    if (transpose == 1 ? (k != 1 || ldw < O) : transpose == 2 ? (ldw < I) : transpose == 3 ? (ldw < O) : (ldw < I * k * k)) return SG3_E_INVALID;
    if (gainMode < 0 || gainMode > 3 || (gainMode && !input_gain)) return SG3_E_INVALID;
    if ((int64_t)N * I > INT32_MAX || (int64_t)I * k * k > INT32_MAX) return SG3_E_TOOLARGE;
    cudaStream_t st = (cudaStream_t)stream;
    if (demodulate) style_norm_kernel<<<1, 1024, 0, st>>>(s, N * I, scratch);
    int gy = N < 8 ? N : 8;
    modconv_weights_kernel<<<dim3((unsigned)O, (unsigned)gy), 256, 0, st>>>(w, s, input_gain, gainMode, wmod, scratch,
                                                                            N, I, O, k * k, ldw, demodulate, round_tf32_flag, transpose);
    return sg3_launch_status(demodulate ? 2 : 1);
}

SG3_EXPORT int sg3_modconv_fwd(const void* x, const float* wmod, void* y,
                               int N, int I, int O, int H, int W, int k, int pad, int ldw,
                               int mathMode, int dtype, void* stream)
{
    return sg3_modconv_fwd_pitched(x, wmod, y, N, I, O, H, W, k, pad, ldw, 0, 0, mathMode, dtype, stream);
}

SG3_EXPORT int sg3_modconv_fwd_pitched(const void* x, const float* wmod, void* y,
                                       int N, int I, int O, int H, int W, int k, int pad, int ldw, int xPitch, int yPitch,
                                       int mathMode, int dtype, void* stream)
{
    if (!x || !wmod || !y || N < 1 || I < 1 || O < 1 || H < 1 || W < 1 || k < 1 || pad < 0) return SG3_E_INVALID;
    const bool tapMajor = mathMode == 1 && k > 1;             // the tensor-core kernels for k > 1 read tap-major weights
    if (ldw < (tapMajor ? I : I * k * k)) return SG3_E_INVALID;
    if (dtype == SG3_F16) {      // fp16 activations and fp16 weights (prologue format 2; tap-major for k = 3): tensor cores only
        if (mathMode != 1) return SG3_E_NOKERNEL;
        if (k == 3) return sg3_modconv_fwd_tc3_f16(x, wmod, y, N, I, O, H, W, pad, ldw, xPitch, yPitch, (cudaStream_t)stream);
        if (yPitch != 0 || xPitch != 0) return SG3_E_NOKERNEL;
        return sg3_modconv_fwd_tc_f16(x, wmod, y, N, I, O, H, W, k, pad, ldw, (cudaStream_t)stream);
    }
    if (dtype != SG3_F32) return SG3_E_NOKERNEL;
    const int OH = H + 2 * pad - k + 1, OW = W + 2 * pad - k + 1;
    if (OH < 1 || OW < 1) return SG3_E_INVALID;
    if ((int64_t)OH * OW > INT32_MAX || (int64_t)I * k * k > INT32_MAX) return SG3_E_TOOLARGE;
    cudaStream_t st = (cudaStream_t)stream;
    if (mathMode == 2)          // 3xTF32: 1x1 kernels (the 3x3 tensor-core kernel has no split variant yet)
        return k == 1 ? sg3_modconv_fwd_tc_x3((const float*)x, wmod, (float*)y, N, I, O, H, W, k, pad, ldw, st) : SG3_E_NOKERNEL;
    if (mathMode == 1 && k == 3) return sg3_modconv_fwd_tc3((const float*)x, wmod, (float*)y, N, I, O, H, W, pad, ldw, xPitch, yPitch, st);
    // only the 3x3 tensor-core kernel reads / writes padded row pitches
    if ((yPitch != 0 && yPitch != OW) || (xPitch != 0 && xPitch != W)) return SG3_E_NOKERNEL;
    if (mathMode == 1) return sg3_modconv_fwd_tc((const float*)x, wmod, (float*)y, N, I, O, H, W, k, pad, ldw, st);
    if (mathMode != 0) return SG3_E_INVALID;
    const int P = OH * OW;
    int64_t total = (int64_t)N * ((O + BO - 1) / BO) * ((P + BP - 1) / BP);
    int64_t cap = (int64_t)sg3_sm_count() * 32;
    unsigned grid = (unsigned)(total < cap ? total : cap);
    modconv_fwd_simt_kernel<<<grid, 256, 0, st>>>((const float*)x, wmod, (float*)y, N, I, O, H, W, k, pad, OH, OW, ldw);
    return sg3_launch_status();
}

SG3_EXPORT int sg3_modconv_weights_bwd(const float* dwmod, const float* w, const float* s, const float* input_gain, int gainMode,
                                       float* dw, float* ds, float* scratch, int N, int I, int O, int ldw, int demodulate, void* stream)
{
    if (!dwmod || !w || !s || !dw || !ds || !scratch || N < 1 || I < 1 || O < 1 || ldw < I) return SG3_E_INVALID;
    if (gainMode < 0 || gainMode > 3 || (gainMode && !input_gain)) return SG3_E_INVALID;
    if (I > 256 * kMaxIPerThread) return SG3_E_NOKERNEL;
    if ((int64_t)N * I > INT32_MAX) return SG3_E_TOOLARGE;
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: [0] = rs, [1 .. 1 + N*I) = dsn accumulator (zeroed here)
    cudaError_t e = cudaMemsetAsync(scratch + 1, 0, (size_t)N * I * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (demodulate) style_norm_kernel<<<1, 1024, 0, st>>>(s, N * I, scratch);
    modconv_weights_bwd_kernel<<<(unsigned)O, 256, 0, st>>>(dwmod, w, s, input_gain, gainMode, scratch, dw, scratch + 1, N, I, O, ldw, demodulate);
    modconv_style_bwd_kernel<<<1, 1024, 0, st>>>(scratch + 1, s, scratch, ds, N * I, demodulate);
    return sg3_launch_status(demodulate ? 3 : 2);
}

SG3_EXPORT int sg3_modconv_tc_supported(int I, int O, int H, int W, int k, int pad)
{
    if (I < 1 || O < 1 || H < 1 || W < 1 || k < 1 || pad < 0) return SG3_E_INVALID;
    if (k == 1) return (pad == 0 && ((int64_t)H * W) % 4 == 0) ? 0 : SG3_E_NOKERNEL;
    return sg3_modconv_tc3_supported(I, O, H, W, k, pad);
}

SG3_EXPORT int sg3_modconv_weights_bwd_taps(const float* dwmod, const float* w, const float* s, const float* input_gain, int gainMode,
                                            float* dw, float* ds, float* scratch, int N, int I, int O, int k, int ldw, int demodulate,
                                            void* stream)
{
    if (!dwmod || !w || !s || !dw || !ds || !scratch || N < 1 || I < 1 || O < 1 || k < 1 || ldw < I) return SG3_E_INVALID;
    if (gainMode < 0 || gainMode > 3 || (gainMode && !input_gain)) return SG3_E_INVALID;
    const int kk = k * k;
    constexpr int kKpt = 18;                                   // 256 x 18 = 4608 = 512 input channels x 9 taps
    if ((int64_t)I * kk > 256 * kKpt) return SG3_E_NOKERNEL;
    if ((int64_t)N * I > INT32_MAX || (int64_t)kk * O * ldw > INT32_MAX) return SG3_E_TOOLARGE;
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: [0] = rs, [1 .. 1 + N*I) = dsn accumulator (zeroed here)
    cudaError_t e = cudaMemsetAsync(scratch + 1, 0, (size_t)N * I * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (demodulate) style_norm_kernel<<<1, 1024, 0, st>>>(s, N * I, scratch);
    modconv_weights_bwd_taps_kernel<kKpt><<<(unsigned)O, 256, 0, st>>>(dwmod, w, s, input_gain, gainMode, scratch, dw, scratch + 1,
                                                                         N, I, O, kk, ldw, demodulate);
    modconv_style_bwd_kernel<<<1, 1024, 0, st>>>(scratch + 1, s, scratch, ds, N * I, demodulate);
    return sg3_launch_status(demodulate ? 3 : 2);
}

SG3_EXPORT int sg3_modconv_wgrad(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int ldw, void* stream)
{
    if (!dy || !x || !dw || N < 1 || I < 1 || O < 1 || H < 1 || W < 1 || ldw < I) return SG3_E_INVALID;
    return sg3_modconv_wgrad_tc(dy, x, dw, N, I, O, H, W, ldw, (cudaStream_t)stream);
}

SG3_EXPORT int sg3_modconv_wgrad3(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int pad, int ldw,
                                  int dyPitch, int xPitch, void* stream)
{
    if (!dy || !x || !dw || N < 1 || I < 1 || O < 1 || H < 1 || W < 1 || pad < 0 || ldw < I || dyPitch < 0 || xPitch < 0) return SG3_E_INVALID;
    return sg3_modconv_wgrad3_tc(dy, x, dw, N, I, O, H, W, pad, ldw, dyPitch, xPitch, (cudaStream_t)stream);
}
