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
    for (int a = 0; a < fH; a++)
        for (int b = 0; b < fW; b++) {
            int sa = flip ? a : fH - 1 - a, sb = flip ? b : fW - 1 - b;
            p.taps[a * fW + b] = f ? f[sa * fW + sb] : 1.0f;
        }
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case SG3_F32: return launch<float>(p, st);
    case SG3_F16: return launch<__half>(p, st);
    case SG3_F64: return launch<double>(p, st);
    }
    return SG3_E_INVALID;
}
