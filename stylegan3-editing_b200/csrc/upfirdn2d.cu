// upfirdn2d.cu -- pad / zero-insert upsample / dense 2-D FIR / decimate, any factors.
//
// Semantics: torch_utils/ops/upfirdn2d.py:168-212 (and the plugin's large-filter gather,
// upfirdn2d.cu:29-92): y[oy][ox] = gain * sum_{a,b} F'[a][b] * Z[oy*downy + a][ox*downx + b]
// with Z the zero-inserted, padded/cropped input and F' the filter flipped unless `flip`.
// Only the polyphase taps that land on real samples are visited.  The taps travel in the
// launch parameters (constant bank) -- no device-global filter state.  This is the generic
// resampler behind the public upfirdn2d op and the fallback composition of filtered_lrelu;
// the synthesis hot path itself runs the fused kernel in filtered_lrelu.cu.
#include "common.cuh"

namespace {

constexpr int kMaxTaps = 768;

struct UpfirdnParams {
    const void* x; void* y;
    int N, C, inH, inW, outH, outW;
    int64_t xs[4], ys[4];             // element strides n, c, h, w
    int fW, fH, upx, upy, downx, downy, padx0, pady0;
    float gain;
    int tilesX, tilesY;
    float taps[kMaxTaps];             // correlation-ordered F'[a][b]
};

constexpr int kTileW = 32, kTileH = 8;

template <class T>
__global__ void __launch_bounds__(kTileW * kTileH) upfirdn2d_kernel(const __grid_constant__ UpfirdnParams p)
{
    typedef typename Arith<T>::type S;
    const int64_t tilesPerPlane = (int64_t)p.tilesX * p.tilesY;
    const int64_t total = tilesPerPlane * p.N * p.C;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int64_t plane = t / tilesPerPlane;
        const int rem = (int)(t - plane * tilesPerPlane);
        const int ty = rem / p.tilesX, tx = rem - ty * p.tilesX;
        const int n = (int)(plane / p.C), c = (int)(plane - (int64_t)n * p.C);
        const int ox = tx * kTileW + (threadIdx.x % kTileW);
        const int oy = ty * kTileH + (threadIdx.x / kTileW);
        if (ox >= p.outW || oy >= p.outH) continue;

        const int midX = ox * p.downx - p.padx0;
        const int midY = oy * p.downy - p.pady0;
        const int b0 = pos_mod(-midX, p.upx);
        const int a0 = pos_mod(-midY, p.upy);
        const int j0 = (midX + b0) / p.upx;       // exact division (may be negative)
        const int i0 = (midY + a0) / p.upy;
        const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];

        S acc = (S)0;
        for (int a = a0, i = i0; a < p.fH; a += p.upy, i++) {
            if (i < 0 || i >= p.inH) continue;
            const T* xr = xp + i * p.xs[2];
            const float* fr = p.taps + a * p.fW;
            for (int b = b0, j = j0; b < p.fW; b += p.upx, j++) {
                if (j < 0 || j >= p.inW) continue;
                acc += ld_as<T>(xr + j * p.xs[3]) * (S)fr[b];
            }
        }
        acc *= (S)p.gain;
        st_as<T>((T*)p.y + n * p.ys[0] + c * p.ys[1] + oy * p.ys[2] + ox * p.ys[3], acc);
    }
}

// ---- fast path: 1-D filters (one axis of a separable filter, the two-launch form of upfirdn2d.py:241-246) ------------
// Compile-time up / down factor and taps per polyphase branch (KP, zero padded), W-contiguous tensors.  Lanes run along
// x so every load and store instruction of a warp is one contiguous segment; the taps of a lane's polyphase branch sit
// in registers (AXIS = x: the branch depends on the lane) or come from the constant bank with a uniform index
// (AXIS = y: the branch depends on the row).  Each thread produces Q outputs so several loads are in flight; the input
// is re-read through L1 (every element is needed by KP outputs), HBM sees each byte once.
constexpr int kFastQ = 4;

template <class T, int AXIS, int UP, int DOWN, int KP>
__global__ void __launch_bounds__(256) upfirdn1d_kernel(const __grid_constant__ UpfirdnParams p)
{
    typedef typename Arith<T>::type S;
    constexpr int TW = AXIS == 0 ? 32 * kFastQ : 32, TH = AXIS == 0 ? 8 : 8 * kFastQ;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int tilesX = (p.outW + TW - 1) / TW, tilesY = (p.outH + TH - 1) / TH;
    const int64_t tilesPerPlane = (int64_t)tilesX * tilesY;
    const int64_t total = tilesPerPlane * p.N * p.C;
    const int pad = AXIS == 0 ? p.padx0 : p.pady0, padOther = AXIS == 0 ? p.pady0 : p.padx0;
    const int inL = AXIS == 0 ? p.inW : p.inH;                   // input extent along the filtered axis
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int64_t plane = t / tilesPerPlane;
        const int rem = (int)(t - plane * tilesPerPlane);
        const int ty = rem / tilesX, tx = rem - ty * tilesX;
        const int n = (int)(plane / p.C), c = (int)(plane - (int64_t)n * p.C);
        const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];
        T* yp = (T*)p.y + n * p.ys[0] + c * p.ys[1];
        if (AXIS == 0) {
            const int oy = ty * TH + wy;
            const int i = oy - padOther;
            if (oy >= p.outH) continue;
            const bool rowOk = i >= 0 && i < p.inH;
            const T* xr = xp + (int64_t)i * p.xs[2];
            // polyphase branch of this lane: the same for all its Q outputs (they are 32 * DOWN apart, a multiple of UP)
            const int ox0 = tx * TW + lane;
            const int a0 = (-(ox0 * DOWN - pad)) & (UP - 1);
            S tap[KP];
#pragma unroll
            for (int k = 0; k < KP; k++) tap[k] = (S)p.taps[a0 + k * UP];
#pragma unroll
            for (int q = 0; q < kFastQ; q++) {
                const int ox = ox0 + 32 * q;
                if (ox >= p.outW) break;
                const int j0 = (ox * DOWN - pad + a0) >> (UP == 4 ? 2 : UP == 2 ? 1 : 0);      // exact; arithmetic shift for negatives
                S acc = (S)0;
                if (rowOk) {
                    if (j0 >= 0 && j0 + KP <= inL) {
#pragma unroll
                        for (int k = 0; k < KP; k++) acc += ld_as<T>(xr + j0 + k) * tap[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; k++)
                            if (j0 + k >= 0 && j0 + k < inL) acc += ld_as<T>(xr + j0 + k) * tap[k];
                    }
                }
                st_as<T>(yp + (int64_t)oy * p.ys[2] + ox, acc * (S)p.gain);
            }
        } else {
            const int ox = tx * TW + lane;
            const int j = ox - padOther;
            if (ox >= p.outW) continue;
            const bool colOk = j >= 0 && j < p.inW;
            const T* xc = xp + j;
#pragma unroll
            for (int q = 0; q < kFastQ; q++) {
                const int oy = ty * TH + wy + 8 * q;
                if (oy >= p.outH) break;
                const int mid = oy * DOWN - pad;
                const int a0 = (-mid) & (UP - 1);                 // uniform across the warp: constant-bank taps
                const int i0 = (mid + a0) >> (UP == 4 ? 2 : UP == 2 ? 1 : 0);
                S acc = (S)0;
                if (colOk) {
                    if (i0 >= 0 && i0 + KP <= inL) {
#pragma unroll
                        for (int k = 0; k < KP; k++) acc += ld_as<T>(xc + (int64_t)(i0 + k) * p.xs[2]) * (S)p.taps[a0 + k * UP];
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; k++)
                            if (i0 + k >= 0 && i0 + k < inL) acc += ld_as<T>(xc + (int64_t)(i0 + k) * p.xs[2]) * (S)p.taps[a0 + k * UP];
                    }
                }
                st_as<T>(yp + (int64_t)oy * p.ys[2] + ox, acc * (S)p.gain);
            }
        }
    }
}

template <class T, int AXIS, int UP, int DOWN, int KP>
int launch_fast(const UpfirdnParams& p, cudaStream_t stream)
{
    const int TW = AXIS == 0 ? 32 * kFastQ : 32, TH = AXIS == 0 ? 8 : 8 * kFastQ;
    const int64_t total = (int64_t)((p.outW + TW - 1) / TW) * ((p.outH + TH - 1) / TH) * p.N * p.C;
    const int64_t cap = (int64_t)sg3_sm_count() * 32;
    upfirdn1d_kernel<T, AXIS, UP, DOWN, KP><<<(unsigned)(total < cap ? total : cap), 256, 0, stream>>>(p);
    return sg3_launch_status();
}

// ---- fast path: small dense filters (the 4x4 outer product that setup_filter builds for [1,3,3,1]: filter2d / upsample2d /
// downsample2d of upfirdn2d.py:278-388), same up / down factor on both axes, at most KP x KP taps per polyphase branch ----
template <class T, int UP, int DOWN, int KP>
__global__ void __launch_bounds__(256) upfirdn2d_small_kernel(const __grid_constant__ UpfirdnParams p)
{
    typedef typename Arith<T>::type S;
    constexpr int TW = 32, TH = 8 * kFastQ, SH = UP == 4 ? 2 : UP == 2 ? 1 : 0;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int tilesX = (p.outW + TW - 1) / TW, tilesY = (p.outH + TH - 1) / TH;
    const int64_t tilesPerPlane = (int64_t)tilesX * tilesY;
    const int64_t total = tilesPerPlane * p.N * p.C;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int64_t plane = t / tilesPerPlane;
        const int rem = (int)(t - plane * tilesPerPlane);
        const int ty = rem / tilesX, tx = rem - ty * tilesX;
        const int n = (int)(plane / p.C), c = (int)(plane - (int64_t)n * p.C);
        const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];
        T* yp = (T*)p.y + n * p.ys[0] + c * p.ys[1];
        const int ox = tx * TW + lane;
        if (ox >= p.outW) continue;
        const int midX = ox * DOWN - p.padx0;
        const int b0 = (-midX) & (UP - 1);
        const int j0 = (midX + b0) >> SH;
        const bool colsIn = j0 >= 0 && j0 + KP <= p.inW;
#pragma unroll
        for (int q = 0; q < kFastQ; q++) {
            const int oy = ty * TH + wy + 8 * q;
            if (oy >= p.outH) break;
            const int midY = oy * DOWN - p.pady0;
            const int a0 = (-midY) & (UP - 1);
            const int i0 = (midY + a0) >> SH;
            S acc = (S)0;
            // taps: 16 x 16 table (launch_small), zero beyond the filter -> no tap predicates
            const float* f0 = p.taps + a0 * 16 + b0;
            const T* x0 = xp + (int64_t)i0 * p.xs[2] + j0;
            if (colsIn && i0 >= 0 && i0 + KP <= p.inH) {          // interior: KP x KP unconditional loads
#pragma unroll
                for (int ka = 0; ka < KP; ka++)
#pragma unroll
                    for (int kb = 0; kb < KP; kb++)
                        acc += ld_as<T>(x0 + (int64_t)ka * p.xs[2] + kb) * (S)f0[ka * UP * 16 + kb * UP];
            } else {
#pragma unroll
                for (int ka = 0; ka < KP; ka++) {
                    const int i = i0 + ka;
                    if (i < 0 || i >= p.inH) continue;
#pragma unroll
                    for (int kb = 0; kb < KP; kb++)
                        if (j0 + kb >= 0 && j0 + kb < p.inW) acc += ld_as<T>(x0 + (int64_t)ka * p.xs[2] + kb) * (S)f0[ka * UP * 16 + kb * UP];
                }
            }
            st_as<T>(yp + (int64_t)oy * p.ys[2] + ox, acc * (S)p.gain);
        }
    }
}

template <class T, int UP, int DOWN, int KP>
int launch_small(const UpfirdnParams& p0, cudaStream_t stream)
{
    UpfirdnParams p = p0;                      // re-pitch the taps to 16 x 16, zero padded (UP * KP <= 16 per axis)
    for (int q = 0; q < 256; q++) p.taps[q] = 0.f;
    for (int a = 0; a < p0.fH; a++)
        for (int b = 0; b < p0.fW; b++) p.taps[a * 16 + b] = p0.taps[a * p0.fW + b];
    const int64_t total = (int64_t)((p.outW + 31) / 32) * ((p.outH + 8 * kFastQ - 1) / (8 * kFastQ)) * p.N * p.C;
    const int64_t cap = (int64_t)sg3_sm_count() * 32;
    upfirdn2d_small_kernel<T, UP, DOWN, KP><<<(unsigned)(total < cap ? total : cap), 256, 0, stream>>>(p);
    return sg3_launch_status();
}

template <class T>
int try_small(const UpfirdnParams& p, cudaStream_t stream)
{
    if (p.xs[3] != 1 || p.ys[3] != 1 || p.upx != p.upy || p.downx != p.downy) return SG3_E_NOKERNEL;
    if ((int64_t)p.outW * 4 > INT32_MAX || (int64_t)p.outH * 4 > INT32_MAX) return SG3_E_NOKERNEL;
    const int up = p.upx, down = p.downx;
    const int kp = ((p.fW > p.fH ? p.fW : p.fH) + up - 1) / up;
    if (kp > 4 || (p.fH + up * 4) * p.fW + up * 4 > kMaxTaps) return SG3_E_NOKERNEL;
#define SG3_SMALL(U, D) if (up == U && down == D) return kp <= 2 ? launch_small<T, U, D, 2>(p, stream) : launch_small<T, U, D, 4>(p, stream);
    SG3_SMALL(1, 1) SG3_SMALL(2, 1) SG3_SMALL(4, 1) SG3_SMALL(1, 2) SG3_SMALL(1, 4)
#undef SG3_SMALL
    return SG3_E_NOKERNEL;
}

template <class T, int AXIS, int UP, int DOWN>
int dispatch_fast_kp(const UpfirdnParams& p, int kp, cudaStream_t stream)
{
    if (kp <= 2) return launch_fast<T, AXIS, UP, DOWN, 2>(p, stream);
    if (kp <= 4) return launch_fast<T, AXIS, UP, DOWN, 4>(p, stream);
    if (kp <= 6) return launch_fast<T, AXIS, UP, DOWN, 6>(p, stream);
    if (kp <= 8) return launch_fast<T, AXIS, UP, DOWN, 8>(p, stream);
    if (kp <= 12) return launch_fast<T, AXIS, UP, DOWN, 12>(p, stream);
    if (kp <= 16) return launch_fast<T, AXIS, UP, DOWN, 16>(p, stream);
    if (kp <= 24) return launch_fast<T, AXIS, UP, DOWN, 24>(p, stream);
    return SG3_E_NOKERNEL;
}

// Returns SG3_E_NOKERNEL when the shape is not a 1-D filter with power-of-two factors on W-contiguous fp32 / fp16 tensors.
template <class T>
int try_fast(const UpfirdnParams& p, cudaStream_t stream)
{
    if (p.xs[3] != 1 || p.ys[3] != 1) return SG3_E_NOKERNEL;
    int axis;
    if (p.fH == 1 && p.upy == 1 && p.downy == 1) axis = 0;
    else if (p.fW == 1 && p.upx == 1 && p.downx == 1) axis = 1;
    else return SG3_E_NOKERNEL;
    const int up = axis == 0 ? p.upx : p.upy, down = axis == 0 ? p.downx : p.downy, fl = axis == 0 ? p.fW : p.fH;
    if ((int64_t)p.outW * 4 > INT32_MAX || (int64_t)p.outH * 4 > INT32_MAX) return SG3_E_NOKERNEL;
    const int kp = (fl + up - 1) / up;
    if (up * 24 + 4 > kMaxTaps) return SG3_E_NOKERNEL;
#define SG3_FAST(AX, U, D) if (axis == AX && up == U && down == D) return dispatch_fast_kp<T, AX, U, D>(p, kp, stream);
    SG3_FAST(0, 1, 1) SG3_FAST(0, 2, 1) SG3_FAST(0, 4, 1) SG3_FAST(0, 1, 2) SG3_FAST(0, 1, 4)
    SG3_FAST(1, 1, 1) SG3_FAST(1, 2, 1) SG3_FAST(1, 4, 1) SG3_FAST(1, 1, 2) SG3_FAST(1, 1, 4)
#undef SG3_FAST
    return SG3_E_NOKERNEL;
}

template <class T>
int launch(const UpfirdnParams& p, cudaStream_t stream)
{
    int64_t total = (int64_t)p.tilesX * p.tilesY * p.N * p.C;
    int64_t cap = (int64_t)sg3_sm_count() * 64;
    unsigned grid = (unsigned)(total < cap ? total : cap);
    upfirdn2d_kernel<T><<<grid, kTileW * kTileH, 0, stream>>>(p);
    return sg3_launch_status();
}

}  // namespace

SG3_EXPORT int sg3_upfirdn2d(const void* x, void* y, const float* f,
                             int N, int C, int inH, int inW, int outH, int outW,
                             const int64_t xStride[4], const int64_t yStride[4],
                             int fW, int fH, int upx, int upy, int downx, int downy,
                             int padx0, int pady0, int flip, float gain,
                             int dtype, void* stream)
{
    if (!x || !y || !xStride || !yStride) return SG3_E_INVALID;
    if (N < 1 || C < 1 || inH < 1 || inW < 1 || outH < 1 || outW < 1) return SG3_E_INVALID;
    if (fW < 1 || fH < 1 || upx < 1 || upy < 1 || downx < 1 || downy < 1) return SG3_E_INVALID;
    if ((int64_t)fW * fH > kMaxTaps) return SG3_E_TOOLARGE;
    UpfirdnParams p;
    p.x = x; p.y = y;
    p.N = N; p.C = C; p.inH = inH; p.inW = inW; p.outH = outH; p.outW = outW;
    for (int i = 0; i < 4; i++) { p.xs[i] = xStride[i]; p.ys[i] = yStride[i]; }
    p.fW = fW; p.fH = fH; p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy;
    p.padx0 = padx0; p.pady0 = pady0; p.gain = gain;
    p.tilesX = (outW + kTileW - 1) / kTileW;
    p.tilesY = (outH + kTileH - 1) / kTileH;
    for (int q = 0; q < kMaxTaps; q++) p.taps[q] = 0.f;             // the 1-D fast path reads zero-padded branches
    for (int a = 0; a < fH; a++)
        for (int b = 0; b < fW; b++) {
            int sa = flip ? a : fH - 1 - a, sb = flip ? b : fW - 1 - b;
            p.taps[a * fW + b] = f ? f[sa * fW + sb] : 1.0f;
        }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SG3_F32 || dtype == SG3_F16) {
        int rc = dtype == SG3_F32 ? try_fast<float>(p, st) : try_fast<__half>(p, st);
        if (rc != SG3_E_NOKERNEL) return rc;
        rc = dtype == SG3_F32 ? try_small<float>(p, st) : try_small<__half>(p, st);
        if (rc != SG3_E_NOKERNEL) return rc;
    }
    switch (dtype) {
    case SG3_F32: return launch<float>(p, st);
    case SG3_F16: return launch<__half>(p, st);
    case SG3_F64: return launch<double>(p, st);
    }
    return SG3_E_INVALID;
}
