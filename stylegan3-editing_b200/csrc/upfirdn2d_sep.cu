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
#include <cstdlib>
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

// Polyphase geometry of a run of consecutive outputs o = base + u whose base is a multiple of UP: with
// rho = pad & (UP - 1), output u uses taps a0(u) + k * UP and inputs first(base) + d(u) + k.  Everything is a compile-time
// function of (u, RHO), so both passes keep their windows in registers with static indices.
template <int UP, int DOWN, int RHO> struct Poly {
    static __host__ __device__ constexpr int a0(int u) { return ((RHO - u * DOWN) % UP + UP) % UP; }
    // first(base + u) - first(base) for base * DOWN a multiple of UP:  ((u * DOWN - RHO' + a0(u)) - (a0(0) - RHO')) / UP
    static __host__ __device__ constexpr int d(int u) { return (u * DOWN + a0(u) - a0(0)) / UP; }
};

template <int UP, int DOWN, int KP, int RUN> struct Win {              // window rows / columns a run of RUN outputs reads
    static constexpr int N = ((RUN - 1) * DOWN + UP - 1) / UP + KP + 1;
};

// x pass of one (row, 8-column block) item: lanes run along the rows, so the shared-memory reads of a warp are IWP (odd)
// floats apart -- conflict-free -- and a thread slides a register window along x.
template <int UP, int DOWN, int KP, int RHO, int IWP, int MP>
__device__ __forceinline__ void sep_xrun(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ fx)
{
    constexpr int NW = Win<UP, DOWN, KP, 8>::N;
    float w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = src[k];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KP; k++) acc = fmaf(w[Poly<UP, DOWN, RHO>::d(u) + k], fx[Poly<UP, DOWN, RHO>::a0(u) + k * UP], acc);
        dst[u] = acc;
    }
}

// y pass of RPT consecutive output rows of one column: the window comes from sMid with lanes along x (conflict-free).
template <class T, int UP, int DOWN, int KP, int RHO, int RPT, int MP>
__device__ __forceinline__ void sep_yrun(const float* __restrict__ src, int rowsAvail, T* __restrict__ out, int64_t outStride, int rowsOut,
                                         const float* __restrict__ fy, float gain)
{
    typedef typename Arith<T>::type S;
    constexpr int NW = Win<UP, DOWN, KP, RPT>::N;
    float w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = k < rowsAvail ? src[k * MP] : 0.f;
#pragma unroll
    for (int q = 0; q < RPT; q++) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KP; k++) acc = fmaf(w[Poly<UP, DOWN, RHO>::d(q) + k], fy[Poly<UP, DOWN, RHO>::a0(q) + k * UP], acc);
        if (q < rowsOut) st_as<T>(out + q * outStride, (S)(acc * gain));
    }
}

template <class T, int UP, int DOWN, int KP>
__global__ void __launch_bounds__(256) upfirdn2d_sep_kernel(const __grid_constant__ SepParams p)
{
    typedef SepGeo<UP, DOWN> G;
    constexpr int TW = G::TW, TH = G::TH;
    constexpr int IW = ((TW - 1) * DOWN + UP - 1) / UP + KP + 1;        // input columns / rows a tile can touch
    constexpr int IH = ((TH - 1) * DOWN + UP - 1) / UP + KP + 1;
    constexpr int IWP = (IW + 8) | 1;                                   // odd pitch (+8: the last x window may run past IW)
    constexpr int MP = TW + 1;                                          // odd pitch of the x-filtered tile
    constexpr int RPT = TH / (256 / TW);                                // output rows per thread in the y pass
    extern __shared__ float smem[];
    float* sIn = smem;                      // [IH][IWP]
    float* sMid = smem + IH * IWP;          // [IH][MP]
    const int tid = threadIdx.x;
    const int64_t tilesPerPlane = (int64_t)p.tilesX * p.tilesY;
    const int64_t total = tilesPerPlane * p.N * p.C;
    const int rhoX = p.padx0 & (UP - 1), rhoY = p.pady0 & (UP - 1);
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
        // ---- input tile -> smem (zero outside the image and in the pitch padding); 8 independent loads per thread in flight ----
        constexpr int NIN = IH * IWP;
        for (int e0 = tid; e0 < NIN; e0 += 256 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = e0 + 256 * u;
                const int r = e / IWP, q = e - r * IWP;
                const int i = i0 + r, j = j0 + q;
                v[u] = 0.f;
                if (e < NIN && q < IW && i >= 0 && i < p.inH && j >= 0 && j < p.inW) v[u] = (float)ld_as<T>(xp + (int64_t)i * p.xs[2] + j);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = e0 + 256 * u;
                if (e < NIN) sIn[e] = v[u];
            }
        }
        __syncthreads();
        // ---- x pass: items (row r, block of 8 output columns), rows fastest so that a warp reads IWP-strided addresses ----
        for (int it = tid; it < IH * (TW / 8); it += 256) {
            const int cb = it / IH, r = it - cb * IH;
            const float* src = sIn + r * IWP + (8 * cb * DOWN) / UP;       // first(ox0 + 8 cb) - j0: 8 cb DOWN is a multiple of UP
            float* dst = sMid + r * MP + 8 * cb;
            if (UP == 1) sep_xrun<UP, DOWN, KP, 0, IWP, MP>(src, dst, p.fx);
            else if (UP == 2) { if (rhoX == 0) sep_xrun<UP, DOWN, KP, 0, IWP, MP>(src, dst, p.fx); else sep_xrun<UP, DOWN, KP, 1 % UP, IWP, MP>(src, dst, p.fx); }
            else {
                if (rhoX == 0) sep_xrun<UP, DOWN, KP, 0, IWP, MP>(src, dst, p.fx);
                else if (rhoX == 1) sep_xrun<UP, DOWN, KP, 1 % UP, IWP, MP>(src, dst, p.fx);
                else if (rhoX == 2) sep_xrun<UP, DOWN, KP, 2 % UP, IWP, MP>(src, dst, p.fx);
                else sep_xrun<UP, DOWN, KP, 3 % UP, IWP, MP>(src, dst, p.fx);
            }
        }
        __syncthreads();
        // ---- y pass: thread = (column, RPT consecutive output rows) ----
        {
            const int cx = tid % TW, rg = tid / TW;
            const int ox = ox0 + cx, oyA = oy0 + rg * RPT;
            if (ox < p.outW && oyA < p.outH) {
                const int rowA = (rg * RPT * DOWN) / UP;                  // first(oyA) - i0 (rg RPT DOWN is a multiple of UP)
                const float* src = sMid + rowA * MP + cx;
                T* out = yp + (int64_t)oyA * p.ys[2] + ox;
                const int rowsOut = min(RPT, p.outH - oyA), rowsAvail = IH - rowA;
                if (UP == 1) sep_yrun<T, UP, DOWN, KP, 0, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain);
                else if (UP == 2) { if (rhoY == 0) sep_yrun<T, UP, DOWN, KP, 0, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain); else sep_yrun<T, UP, DOWN, KP, 1 % UP, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain); }
                else {
                    if (rhoY == 0) sep_yrun<T, UP, DOWN, KP, 0, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain);
                    else if (rhoY == 1) sep_yrun<T, UP, DOWN, KP, 1 % UP, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain);
                    else if (rhoY == 2) sep_yrun<T, UP, DOWN, KP, 2 % UP, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain);
                    else sep_yrun<T, UP, DOWN, KP, 3 % UP, RPT, MP>(src, rowsAvail, out, p.ys[2], rowsOut, p.fy, p.gain);
                }
            }
        }
    }
}

template <class T, int UP, int DOWN, int KP>
int launch_sep(const SepParams& p0, cudaStream_t stream)
{
    typedef SepGeo<UP, DOWN> G;
    constexpr int IW = ((G::TW - 1) * DOWN + UP - 1) / UP + KP + 1, IH = ((G::TH - 1) * DOWN + UP - 1) / UP + KP + 1;
    constexpr int smemBytes = (IH * ((IW + 8) | 1) + IH * (G::TW + 1)) * 4;
    SepParams p = p0;
    p.tilesX = (p.outW + G::TW - 1) / G::TW;
    p.tilesY = (p.outH + G::TH - 1) / G::TH;
    auto kern = upfirdn2d_sep_kernel<T, UP, DOWN, KP>;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([&] {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
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

int sg3_upfirdn2d_stream(const float* x, float* y, int N, int C, int inH, int inW, int outH, int outW,
                         const int64_t xs[4], const int64_t ys[4], const float* fx, const float* fy, int fW, int fH,
                         int up, int down, int padx0, int pady0, float gain, cudaStream_t stream);

// SG3_UPFIRDN_TILED=1 in the environment keeps the tiled kernel for everything (A/B timing, tools/prof_ops.py)
static bool stream_kernel_disabled()
{
    static const bool off = [] { const char* e = getenv("SG3_UPFIRDN_TILED"); return e && e[0] == '1'; }();
    return off;
}

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
    if (dtype == SG3_F32 && !stream_kernel_disabled()) {
        // warp-streaming kernel first (upfirdn2d_stream.cu: up 2 / down 2 / neither, <= 12 taps per branch); the tiled kernel
        // below covers the rest (factors of 4, longer filters, fp16)
        const int rc = sg3_upfirdn2d_stream((const float*)x, (float*)y, N, C, inH, inW, outH, outW, xStride, yStride, p.fx, p.fy, fW, fH,
                                            up, down, padx0, pady0, gain, st);
        if (rc != SG3_E_NOKERNEL) return rc;
    }
    return dtype == SG3_F32 ? dispatch_sep<float>(p, up, down, st) : dispatch_sep<__half>(p, up, down, st);
}
