// flrelu_act.cu -- in-place x = clamp(lrelu(x * gain)) with the 2-bit sign tensor of the
// fused op, for the generic (unfused) composition of filtered_lrelu.
//
// Semantics: torch_utils/ops/filtered_lrelu.cu:1105-1211.  Sign codes: 1 = value was
// negative (scaled by slope), 2 = clamped (gradient zero); 4 pixels per byte along x,
// pixel k of a byte at bits 2k..2k+1, rows padded to a multiple of 16 pixels.
//   write: codes stored at (x, y) of the data tensor (offsets are not applied, like the reference)
//   read : v = x*gain * {1, slope, 0}[code at (x+sx, y+sy)], unchanged outside the sign tensor
#include "common.cuh"

namespace {

struct ActParams {
    void* x; uint8_t* s;
    int N, C, H, W;
    int64_t xs[4];          // element strides
    int sH, sWb, sx, sy;
    float gain, slope, clamp;
};

// One thread = 4 horizontally adjacent pixels = one sign byte; a warp covers 128 pixels of a row.
template <class T, int MODE>
__global__ void __launch_bounds__(256) flrelu_act_kernel(const __grid_constant__ ActParams p)
{
    typedef typename Arith<T>::type S;
    const int Wq = MODE == SG3_SIGNS_WRITE ? p.sWb : (p.W + 3) >> 2;     // quads per row
    const int rows = MODE == SG3_SIGNS_WRITE ? p.sH : p.H;
    const int64_t perPlane = (int64_t)Wq * rows;
    const int64_t total = perPlane * p.N * p.C;
    const S gain = (S)p.gain, slope = (S)p.slope, clamp = (S)p.clamp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t plane = idx / perPlane;
        const int64_t rem = idx - plane * perPlane;
        const int yy = (int)(rem / Wq), q = (int)(rem - (int64_t)yy * Wq);
        const int n = (int)(plane / p.C), c = (int)(plane - (int64_t)n * p.C);
        T* row = (T*)p.x + n * p.xs[0] + c * p.xs[1] + yy * p.xs[2];
        unsigned code = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = q * 4 + k;
            if (xx >= p.W || yy >= p.H) continue;
            T* pv = row + xx * p.xs[3];
            S v = ld_as<T>(pv) * gain;
            if (MODE == SG3_SIGNS_READ) {
                const int px = xx + p.sx, py = yy + p.sy;
                if (px >= 0 && py >= 0 && py < p.sH && (px >> 2) < p.sWb) {
                    const unsigned sb = p.s[((int64_t)plane * p.sH + py) * p.sWb + (px >> 2)] >> ((px & 3) * 2);
                    if (sb & 1) v *= slope;
                    if (sb & 2) v = (S)0;
                }
            } else {
                unsigned cpx = 0;
                if (v < (S)0) { v *= slope; cpx = 1; }
                if (fabs(v) > clamp) { v = v < (S)0 ? -clamp : clamp; cpx = 2; }
                code |= cpx << (2 * k);
            }
            st_as<T>(pv, v);
        }
        if (MODE == SG3_SIGNS_WRITE) p.s[((int64_t)plane * p.sH + yy) * p.sWb + q] = (uint8_t)code;
    }
}

template <class T>
int launch(const ActParams& p, int mode, cudaStream_t stream)
{
    const int Wq = mode == SG3_SIGNS_WRITE ? p.sWb : (p.W + 3) >> 2;
    const int rows = mode == SG3_SIGNS_WRITE ? p.sH : p.H;
    int64_t total = (int64_t)Wq * rows * p.N * p.C;
    int64_t blocks = ceil_div64(total, 256);
    int64_t cap = (int64_t)sg3_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    switch (mode) {
    case SG3_SIGNS_NONE:  flrelu_act_kernel<T, SG3_SIGNS_NONE><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    case SG3_SIGNS_WRITE: flrelu_act_kernel<T, SG3_SIGNS_WRITE><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    case SG3_SIGNS_READ:  flrelu_act_kernel<T, SG3_SIGNS_READ><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    default: return SG3_E_INVALID;
    }
    return sg3_launch_status();
}

// ---- pointwise filtered_lrelu: up = down = 1 with 1x1 filters (the ToRGB layer, networks_stylegan3.py:360-363) --------
// y = fdScale * act(fuScale * gain * (x + b)) with the same sign-tensor semantics as the fused kernel; pure HBM streaming,
// one thread per 4 horizontally adjacent pixels (= one sign byte).
struct PointParams {
    const void* x; void* y; const void* b; uint8_t* s;
    int N, C, H, W;
    long long xs[4], ys[4], bs;       // byte strides
    int sH, sWb, sx, sy;
    float pre, slope, clamp, post;    // pre = fuScale * gain, post = fdScale
};

template <class T, int MODE>
__global__ void __launch_bounds__(256) flrelu_pointwise_kernel(const __grid_constant__ PointParams p)
{
    const int Wq = (p.W + 3) >> 2;
    const long long perPlane = (long long)Wq * p.H;
    const long long total = perPlane * p.N * p.C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long plane = idx / perPlane;
        const long long rem = idx - plane * perPlane;
        const int yy = (int)(rem / Wq), q = (int)(rem - (long long)yy * Wq);
        const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
        const char* xr = (const char*)p.x + n * p.xs[0] + c * p.xs[1] + yy * p.xs[2];
        char* yr = (char*)p.y + n * p.ys[0] + c * p.ys[1] + yy * p.ys[2];
        const float bias = p.b ? (float)ld_as<T>((const T*)((const char*)p.b + c * p.bs)) : 0.f;
        unsigned code = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = q * 4 + k;
            if (xx >= p.W) continue;
            float v = ((float)ld_as<T>((const T*)(xr + xx * p.xs[3])) + bias) * p.pre;
            if (MODE == SG3_SIGNS_READ) {
                const int px = xx + p.sx, py = yy + p.sy;
                if (px >= 0 && py >= 0 && py < p.sH && (px >> 2) < p.sWb) {
                    const unsigned sb = p.s[((long long)plane * p.sH + py) * p.sWb + (px >> 2)] >> ((px & 3) * 2);
                    if (sb & 1) v *= p.slope;
                    if (sb & 2) v = 0.f;
                }
            } else {
                unsigned cpx = 0;
                if (v < 0.f) { v *= p.slope; cpx = 1; }
                if (fabsf(v) > p.clamp) { v = v < 0.f ? -p.clamp : p.clamp; cpx = 2; }
                code |= cpx << (2 * k);
            }
            st_as<T>((T*)(yr + xx * p.ys[3]), v * p.post);
        }
        if (MODE == SG3_SIGNS_WRITE) {
            const int py = yy + p.sy, pb = q + (p.sx >> 2);          // sx is a multiple of 4 (checked by the caller)
            if (py >= 0 && py < p.sH && pb >= 0 && pb < p.sWb) p.s[((long long)plane * p.sH + py) * p.sWb + pb] = (uint8_t)code;
        }
    }
}

template <class T>
int launch_point(const PointParams& p, int mode, cudaStream_t stream)
{
    long long total = (long long)((p.W + 3) >> 2) * p.H * p.N * p.C;
    long long blocks = ceil_div64(total, 256);
    long long cap = (long long)sg3_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    switch (mode) {
    case SG3_SIGNS_NONE:  flrelu_pointwise_kernel<T, SG3_SIGNS_NONE><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    case SG3_SIGNS_WRITE: flrelu_pointwise_kernel<T, SG3_SIGNS_WRITE><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    case SG3_SIGNS_READ:  flrelu_pointwise_kernel<T, SG3_SIGNS_READ><<<(unsigned)blocks, 256, 0, stream>>>(p); break;
    default: return SG3_E_INVALID;
    }
    return sg3_launch_status();
}

}  // namespace

// Called by sg3_filtered_lrelu for up == down == 1 with 1x1 filters and no padding.
int sg3_flrelu_pointwise(const sg3_flrelu_desc* d, float fuScale, float fdScale, cudaStream_t stream)
{
    PointParams p;
    p.x = d->x; p.y = d->y; p.b = d->b; p.s = d->signs;
    p.N = d->N; p.C = d->C; p.H = d->inH; p.W = d->inW;
    for (int i = 0; i < 4; i++) { p.xs[i] = d->xStride[i]; p.ys[i] = d->yStride[i]; }
    p.bs = d->bStride;
    p.sH = d->sH; p.sWb = d->sWb; p.sx = d->sx; p.sy = d->sy;
    p.pre = fuScale * d->gain; p.slope = d->slope; p.clamp = d->clamp; p.post = fdScale;
    if (d->dtype == SG3_F32) return launch_point<float>(p, d->signMode, stream);
    if (d->dtype == SG3_F16) return launch_point<__half>(p, d->signMode, stream);
    return SG3_E_NOKERNEL;
}

SG3_EXPORT int sg3_filtered_lrelu_act(void* x, uint8_t* signs,
                                      int N, int C, int H, int W, const int64_t xStride[4],
                                      int sH, int sWb, int sx, int sy,
                                      float gain, float slope, float clamp,
                                      int signMode, int dtype, void* stream)
{
    if (!x || !xStride || N < 1 || C < 1 || H < 1 || W < 1) return SG3_E_INVALID;
    if (signMode != SG3_SIGNS_NONE && (!signs || sH < 1 || sWb < 1)) return SG3_E_INVALID;
    if (signMode == SG3_SIGNS_WRITE && (sH < H || sWb * 4 < W)) return SG3_E_INVALID;
    ActParams p;
    p.x = x; p.s = signs; p.N = N; p.C = C; p.H = H; p.W = W;
    for (int i = 0; i < 4; i++) p.xs[i] = xStride[i];
    p.sH = sH; p.sWb = sWb; p.sx = sx; p.sy = sy;
    p.gain = gain; p.slope = slope; p.clamp = clamp;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case SG3_F32: return launch<float>(p, signMode, st);
    case SG3_F16: return launch<__half>(p, signMode, st);
    case SG3_F64: return launch<double>(p, signMode, st);
    }
    return SG3_E_INVALID;
}
