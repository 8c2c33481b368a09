// upfirdn2d_sep.cu -- separable upfirdn2d in ONE pass over HBM.
//
// The reference runs a separable filter as two launches (x pass, then y pass: upfirdn2d.py:241-246) and so writes and
// re-reads an intermediate image that is `up` (or 1/down) times the input.  Here a CTA stages the input tile of its
// 64 x 32 (down 4: 32 x 16) output tile in shared memory, filters it along x into a second shared-memory tile and along y
// straight into the output: HBM sees the input and the output once (plus the tile halo).  Same arithmetic as the two
// passes (fp32 accumulation; the intermediate stays fp32 also for fp16 tensors).
//
// Up / down factor (the same on both axes, one of them 1) and the taps per polyphase branch (KP) are compile-time;
// lanes run along x: global loads / stores are contiguous segments, the x pass keeps the taps of a lane's branch in
// registers, the y pass reads its taps with a warp-uniform index.
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kMaxSepTaps = 24 * 4 + 8;      // UP * KP entries per axis, zero padded

struct SepParams {
    const void* x; void* y;
    int N, C, inH, inW, outH, outW;
    int64_t xs[4], ys[4];             // element strides
    int fW, fH, padx0, pady0;
    float gain;
    int tilesX, tilesY;
    float fx[kMaxSepTaps], fy[kMaxSepTaps];      // correlation-ordered taps, zero padded
};

template <int UP, int DOWN> struct SepGeo {
    static constexpr int TW = DOWN == 4 ? 32 : 64, TH = DOWN == 4 ? 16 : 32;
    static constexpr int SH = UP == 4 ? 2 : UP == 2 ? 1 : 0;
};

// first input index and polyphase branch of output index o along one axis
template <int UP, int DOWN>
__device__ __forceinline__ void branch(int o, int pad, int& first, int& a0)
{
    const int mid = o * DOWN - pad;
    a0 = (-mid) & (UP - 1);
    first = (mid + a0) >> SepGeo<UP, DOWN>::SH;      // exact division; arithmetic shift for negatives
}

template <class T, int UP, int DOWN, int KP>
__global__ void __launch_bounds__(256) upfirdn2d_sep_kernel(const __grid_constant__ SepParams p)
{
    typedef SepGeo<UP, DOWN> G;
    constexpr int TW = G::TW, TH = G::TH;
    constexpr int IW = ((TW - 1) * DOWN + UP - 1) / UP + KP + 1;        // input columns / rows a tile can touch
    constexpr int IH = ((TH - 1) * DOWN + UP - 1) / UP + KP + 1;
    constexpr int IWP = IW | 1;                                         // odd pitch: rows start in different banks
    extern __shared__ float smem[];
    float* sIn = smem;                      // [IH][IWP]
    float* sMid = smem + IH * IWP;          // [IH][TW]
    const int tid = threadIdx.x;
    const int64_t tilesPerPlane = (int64_t)p.tilesX * p.tilesY;
    const int64_t total = tilesPerPlane * p.N * p.C;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int64_t plane = t / tilesPerPlane;
        const int rem = (int)(t - plane * tilesPerPlane);
        const int ty = rem / p.tilesX, tx = rem - ty * p.tilesX;
        const int n = (int)(plane / p.C), c = (int)(plane - (int64_t)n * p.C);
        const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];
        T* yp = (T*)p.y + n * p.ys[0] + c * p.ys[1];
        const int ox0 = tx * TW, oy0 = ty * TH;
        int j0, i0, dummy;
        branch<UP, DOWN>(ox0, p.padx0, j0, dummy);
        branch<UP, DOWN>(oy0, p.pady0, i0, dummy);
        __syncthreads();                                    // previous tile's readers are done with both buffers
        // ---- input tile -> smem (zero outside the image); 8 independent loads per thread in flight ----
        constexpr int NIN = IH * IW;
        for (int e0 = tid; e0 < NIN; e0 += 256 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = e0 + 256 * u;
                const int r = e / IW, q = e - r * IW;
                const int i = i0 + r, j = j0 + q;
                v[u] = 0.f;
                if (e < NIN && i >= 0 && i < p.inH && j >= 0 && j < p.inW) v[u] = (float)ld_as<T>(xp + (int64_t)i * p.xs[2] + j);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = e0 + 256 * u;
                const int r = e / IW, q = e - r * IW;
                if (e < NIN) sIn[r * IWP + q] = v[u];
            }
        }
        __syncthreads();
        // ---- x pass: sMid[r][cx] = sum_k fx[b0 + k UP] * sIn[r][first(cx) - j0 + k] ----
        {
            const int cx = tid % TW, rg = tid / TW;             // column of this thread, first row; rows step 256 / TW
            int first, b0;
            branch<UP, DOWN>(ox0 + cx, p.padx0, first, b0);
            float tap[KP];
#pragma unroll
            for (int k = 0; k < KP; k++) tap[k] = p.fx[b0 + k * UP];
            const float* src = sIn + (first - j0);
            for (int r = rg; r < IH; r += 256 / TW) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < KP; k++) acc = fmaf(src[r * IWP + k], tap[k], acc);
                sMid[r * TW + cx] = acc;
            }
        }
        __syncthreads();
        // ---- y pass: y[oy][ox] = gain * sum_k fy[a0 + k UP] * sMid[first(oy) - i0 + k][cx] ----
        {
            const int cx = tid % TW, rg = tid / TW;
            const int ox = ox0 + cx;
            for (int ry = rg; ry < TH; ry += 256 / TW) {
                const int oy = oy0 + ry;
                int first, a0;
                branch<UP, DOWN>(oy, p.pady0, first, a0);      // uniform across the warp (TW >= 32): uniform tap loads
                const float* src = sMid + (first - i0) * TW + cx;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < KP; k++) acc = fmaf(src[k * TW], p.fy[a0 + k * UP], acc);
                if (ox < p.outW && oy < p.outH) st_as<T>(yp + (int64_t)oy * p.ys[2] + ox, acc * p.gain);
            }
        }
    }
}

template <class T, int UP, int DOWN, int KP>
int launch_sep(const SepParams& p0, cudaStream_t stream)
{
    typedef SepGeo<UP, DOWN> G;
    constexpr int IW = ((G::TW - 1) * DOWN + UP - 1) / UP + KP + 1, IH = ((G::TH - 1) * DOWN + UP - 1) / UP + KP + 1;
    constexpr int smemBytes = (IH * (IW | 1) + IH * G::TW) * 4;
    SepParams p = p0;
    p.tilesX = (p.outW + G::TW - 1) / G::TW;
    p.tilesY = (p.outH + G::TH - 1) / G::TH;
    auto kern = upfirdn2d_sep_kernel<T, UP, DOWN, KP>;
    static std::once_flag once;
    static cudaError_t attrErr = cudaSuccess;
    std::call_once(once, [&] {
        attrErr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
        if (attrErr == cudaSuccess) attrErr = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    });
    if (attrErr != cudaSuccess) return (int)attrErr;
    const int64_t total = (int64_t)p.tilesX * p.tilesY * p.N * p.C;
    const int64_t cap = (int64_t)sg3_sm_count() * 16;
    kern<<<(unsigned)(total < cap ? total : cap), 256, smemBytes, stream>>>(p);
    return sg3_launch_status();
}

template <class T, int UP, int DOWN>
int dispatch_sep_kp(const SepParams& p, int kp, cudaStream_t stream)
{
    if (kp <= 4) return launch_sep<T, UP, DOWN, 4>(p, stream);
    if (kp <= 6) return launch_sep<T, UP, DOWN, 6>(p, stream);
    if (kp <= 8) return launch_sep<T, UP, DOWN, 8>(p, stream);
    if (kp <= 12) return launch_sep<T, UP, DOWN, 12>(p, stream);
    if (kp <= 24) return launch_sep<T, UP, DOWN, 24>(p, stream);
    return SG3_E_NOKERNEL;
}

template <class T>
int dispatch_sep(const SepParams& p, int up, int down, cudaStream_t stream)
{
    const int fl = p.fW > p.fH ? p.fW : p.fH;
    const int kp = (fl + up - 1) / up;
#define SG3_SEP(U, D) if (up == U && down == D) return dispatch_sep_kp<T, U, D>(p, kp, stream);
    SG3_SEP(1, 1) SG3_SEP(2, 1) SG3_SEP(4, 1) SG3_SEP(1, 2) SG3_SEP(1, 4)
#undef SG3_SEP
    return SG3_E_NOKERNEL;
}

}  // namespace

// Separable filter, same factors on both axes.  fx [fW], fy [fH] host taps (NULL = single 1); other arguments as sg3_upfirdn2d.
SG3_EXPORT int sg3_upfirdn2d_sep(const void* x, void* y, const float* fx, const float* fy,
                                 int N, int C, int inH, int inW, int outH, int outW,
                                 const int64_t xStride[4], const int64_t yStride[4],
                                 int fW, int fH, int up, int down, int padx0, int pady0, int flip, float gain,
                                 int dtype, void* stream)
{
    if (!x || !y || !xStride || !yStride) return SG3_E_INVALID;
    if (N < 1 || C < 1 || inH < 1 || inW < 1 || outH < 1 || outW < 1 || fW < 1 || fH < 1 || up < 1 || down < 1) return SG3_E_INVALID;
    if (dtype != SG3_F32 && dtype != SG3_F16) return SG3_E_NOKERNEL;
    if (xStride[3] != 1 || yStride[3] != 1) return SG3_E_NOKERNEL;
    if ((up != 1 && down != 1) || (int64_t)outW * 4 > INT32_MAX || (int64_t)outH * 4 > INT32_MAX) return SG3_E_NOKERNEL;
    if (fW > 24 * up || fH > 24 * up) return SG3_E_NOKERNEL;
    SepParams p;
    p.x = x; p.y = y; p.N = N; p.C = C; p.inH = inH; p.inW = inW; p.outH = outH; p.outW = outW;
    for (int i = 0; i < 4; i++) { p.xs[i] = xStride[i]; p.ys[i] = yStride[i]; }
    p.fW = fW; p.fH = fH; p.padx0 = padx0; p.pady0 = pady0; p.gain = gain;
    for (int q = 0; q < kMaxSepTaps; q++) { p.fx[q] = 0.f; p.fy[q] = 0.f; }
    for (int b = 0; b < fW; b++) p.fx[b] = fx ? fx[flip ? b : fW - 1 - b] : 1.0f;
    for (int a = 0; a < fH; a++) p.fy[a] = fy ? fy[flip ? a : fH - 1 - a] : 1.0f;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == SG3_F32 ? dispatch_sep<float>(p, up, down, st) : dispatch_sep<__half>(p, up, down, st);
}
