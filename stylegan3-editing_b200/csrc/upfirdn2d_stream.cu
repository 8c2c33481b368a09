// upfirdn2d_stream.cu -- separable upfirdn2d (fp32; up 2, down 2 or neither; <= 12 taps per polyphase branch) as a
// warp-streaming kernel: HBM sees the input once and the output once, nothing else.
//
// The reference runs a separable filter as two launches through an intermediate image (upfirdn2d.py:241-246, kernels of
// upfirdn2d.cu:29-375); upfirdn2d_sep.cu fuses the two passes in a CTA tile but pays two __syncthreads, scalar shared-memory
// traffic and a 27 % tile halo, and ends at 0.3 of the HBM roof.  This kernel follows flrelu_stream instead:
//   * one WARP owns a strip of 128 output columns (lane = 4 adjacent columns: one 128-bit store per lane and row, a warp
//     writes 512 contiguous bytes) and streams down a chunk of rows; warps never synchronise with each other;
//   * input rows arrive in a per-warp shared-memory ring by cp.async (16-byte copies when the rows allow it, else 8-byte;
//     columns and rows outside the image are zero-filled by the copy itself), five rows in flight per warp;
//   * the x pass reads the lane's window with 128-bit (64-bit for up 2) shared-memory loads and leaves 4 values in registers;
//   * the y pass keeps a sliding window of x-filtered rows in REGISTERS (static slots: the loop body is unrolled over one
//     rotation of the window), so the intermediate image of the reference never exists anywhere.
// All alignment cases are folded into the tap tables on the host: the staged row always starts at a column that is a multiple
// of 4 (aligned copies and loads); the 0-3 columns between that and the first column a strip needs, and the polyphase branch of
// each of a lane's 4 outputs, just shift the taps inside a slightly longer static tap range (zero taps outside the filter), and
// the taps are kernel parameters, i.e. constant-bank operands of the FFMAs.  Same arithmetic as the two passes of the
// reference (fp32 accumulation, gain applied once at the end).
#include <cuda_runtime.h>

#include "common.cuh"

namespace ufs {

constexpr int kWarps = 4;
constexpr int kRing = 6;             // staged input rows per warp (kRing - 1 copies in flight)
constexpr int kV = 4;                // output columns per lane
constexpr int kMaxNT = 16, kMaxWin = 16;

template <int UP, int DOWN, int KP> struct Geo {
    static_assert((UP == 1 || UP == 2) && (DOWN == 1 || DOWN == 2) && !(UP == 2 && DOWN == 2), "factors");
    static constexpr int T = UP * KP;                                    // taps, zero padded
    static constexpr int XSTRIDE = 128 * DOWN / UP;                      // input columns between strips
    static constexpr int LSTEP = kV * DOWN / UP;                         // input columns between lanes (2 / 8 / 4)
    static constexpr int NT = UP == 2 ? KP + 2 : T + 3;                  // static tap range per output (filter + alignment slack)
    static constexpr int wlo(int v) { return UP == 2 ? v / 2 : v * DOWN; }          // first window position output v can touch
    static constexpr int LG = UP == 2 ? 2 : 4;                           // floats per shared-memory load of the window
    static constexpr int NWH = ((wlo(kV - 1) + NT + LG - 1) / LG) * LG;  // lane window (floats)
    static constexpr int NINP = ((31 * LSTEP + (UP == 2 ? 2 : 0) + NWH + 3) / 4) * 4;   // staged floats per row
    static constexpr int WIN = UP == 2 ? KP + 1 : T;                     // y window (rows of x-filtered values)
    static constexpr int RPI = DOWN;                                     // input rows consumed per iteration
    static constexpr int PRIME = WIN - RPI;                              // rows consumed before the first output
    static constexpr int PERIOD = WIN / RPI;                             // iterations per rotation of the window
    static_assert(WIN % RPI == 0 && NT <= kMaxNT && WIN <= kMaxWin, "window geometry");
};

struct Params {
    const float* x; float* y;
    int N, C, inH, inW, outH, outW;
    long long xs0, xs1, xs2, ys0, ys1, ys2;      // element strides (unit stride along x)
    float gain;
    int stripsX, chunksY, chunkRows;
    long long totalWarps;
    int jA0;          // first staged input column of strip 0 (multiple of 4, may be negative)
    int c0;           // offset of lane 0's window inside the staged row (up 2: 0 or 2)
    int iBase0;       // input row at window position 0 of the chunk that starts at output row 0
    int vecStore;     // output rows are 16-byte aligned: one 128-bit store per lane
    float th[kV][kMaxNT];      // x pass: taps of output v over window positions wlo(v) .. wlo(v) + NT - 1
    float tv[2][kMaxWin];      // y pass: taps over the window rows (up 2: one table per output row of a pair)
};

template <int BYTES>
__device__ __forceinline__ void cp_async(float* dst, const float* src, int srcBytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(srcBytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(srcBytes) : "memory");
}

template <int UP, int DOWN, int KP, int COPY>
__global__ void __launch_bounds__(kWarps * 32) kernel(const __grid_constant__ Params p)
{
    typedef Geo<UP, DOWN, KP> G;
    extern __shared__ __align__(16) float smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kWarps + warp;
    if (wid >= p.totalWarps) return;
    const int strip = (int)(wid % p.stripsX);
    const long long rest = wid / p.stripsX;
    const int chunk = (int)(rest % p.chunksY);
    const long long plane = rest / p.chunksY;
    const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
    const float* xp = p.x + n * p.xs0 + c * p.xs1;
    float* yp = p.y + n * p.ys0 + c * p.ys1;
    const int ox = strip * 128 + kV * lane;                 // first of this lane's 4 output columns
    const int oy0 = chunk * p.chunkRows;
    const int chs = min(p.chunkRows, p.outH - oy0);
    const int jA = p.jA0 + strip * G::XSTRIDE;
    const int iBase = p.iBase0 + chunk * (p.chunkRows * DOWN / UP);
    const int nIt = UP == 2 ? (chs + 1) >> 1 : chs;
    const int nRows = G::PRIME + G::RPI * nIt;
    float* ring = smem + warp * (kRing * G::NINP);

    // ---- stage input row iBase + r into ring slot `slot` (one commit group per call, empty past the last row) ----
    // Per lane and 16- / 8-byte piece of the row, everything that does not depend on the row is computed once: the column
    // (clamped into the image so that the source address is always valid and aligned) and the bytes of the piece that lie
    // inside the image; a row outside the image copies 0 bytes (cp.async zero-fills the rest of the piece).
    constexpr int CE = COPY / 4;                                  // floats per piece
    constexpr int NF = (G::NINP / CE + 31) / 32;                  // pieces per lane
    int colC[NF], nbC[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const int e = CE * (lane + 32 * f), j = jA + e;
        nbC[f] = (e < G::NINP && j >= 0) ? min(max(p.inW - j, 0), CE) * 4 : 0;
        colC[f] = nbC[f] ? j : 0;
        if (e >= G::NINP) nbC[f] = -1;                            // this lane has no piece f
    }
    auto issue = [&](int r, int slot) {
        if (r < nRows) {
            const int i = iBase + r;
            const int rowMask = (i >= 0 && i < p.inH) ? -1 : 0;
            const float* src = xp + (long long)min(max(i, 0), p.inH - 1) * p.xs2;
            float* dst = ring + slot * G::NINP + CE * lane;
#pragma unroll
            for (int f = 0; f < NF; f++)
                if (nbC[f] >= 0) cp_async<COPY>(dst + CE * 32 * f, src + colC[f], nbC[f] & rowMask);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int rNext = 0, slotIssue = 0, slotUse = 0;
#pragma unroll 1
    for (; rNext < kRing - 1; rNext++) {
        issue(rNext, slotIssue);
        slotIssue = slotIssue + 1 == kRing ? 0 : slotIssue + 1;
    }

    // ---- x pass of the next staged row: 4 values per lane ----
    auto xrow = [&](float (&h)[kV]) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 2) : "memory");
        __syncwarp();
        const float* row = ring + slotUse * G::NINP + p.c0 + G::LSTEP * lane;
        float xin[G::NWH];
        if (UP == 2) {
#pragma unroll
            for (int w = 0; w < G::NWH; w += 2) {
                const float2 t = *reinterpret_cast<const float2*>(row + w);
                xin[w] = t.x; xin[w + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int w = 0; w < G::NWH; w += 4) {
                const float4 t = *reinterpret_cast<const float4*>(row + w);
                xin[w] = t.x; xin[w + 1] = t.y; xin[w + 2] = t.z; xin[w + 3] = t.w;
            }
        }
        slotUse = slotUse + 1 == kRing ? 0 : slotUse + 1;
        // the slot read one row ago is free now (every lane has passed the __syncwarp above): refill it
        issue(rNext, slotIssue);
        rNext++;
        slotIssue = slotIssue + 1 == kRing ? 0 : slotIssue + 1;
#pragma unroll
        for (int v = 0; v < kV; v++) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < G::NT; t++) acc = fmaf(xin[G::wlo(v) + t], p.th[v][t], acc);
            h[v] = acc;
        }
    };

    auto store = [&](int oy, const float (&o)[kV]) {
        if (oy >= p.outH || ox >= p.outW) return;
        float* q = yp + (long long)oy * p.ys2 + ox;
        if (p.vecStore && ox + 3 < p.outW) {
            *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int v = 0; v < kV; v++)
                if (ox + v < p.outW) q[v] = o[v];
        }
    };

    float win[G::WIN][kV];
#pragma unroll
    for (int r = 0; r < G::PRIME; r++) xrow(win[r]);

#pragma unroll 1
    for (int m0 = 0; m0 < nIt; m0 += G::PERIOD) {
#pragma unroll
        for (int q = 0; q < G::PERIOD; q++) {
            const int m = m0 + q;
            if (m < nIt) {
#pragma unroll
                for (int j = 0; j < G::RPI; j++) xrow(win[(G::RPI * q + G::PRIME + j) % G::WIN]);
                if (UP == 2) {
                    float a[kV], b[kV];
#pragma unroll
                    for (int v = 0; v < kV; v++) { a[v] = 0.f; b[v] = 0.f; }
#pragma unroll
                    for (int w = 0; w < G::WIN; w++)
#pragma unroll
                        for (int v = 0; v < kV; v++) {
                            const float hv = win[(q + w) % G::WIN][v];
                            a[v] = fmaf(hv, p.tv[0][w], a[v]);
                            b[v] = fmaf(hv, p.tv[1][w], b[v]);
                        }
#pragma unroll
                    for (int v = 0; v < kV; v++) { a[v] *= p.gain; b[v] *= p.gain; }
                    store(oy0 + 2 * m, a);
                    if (2 * m + 1 < chs) store(oy0 + 2 * m + 1, b);
                } else {
                    float a[kV];
#pragma unroll
                    for (int v = 0; v < kV; v++) a[v] = 0.f;
#pragma unroll
                    for (int w = 0; w < G::WIN; w++)
#pragma unroll
                        for (int v = 0; v < kV; v++) a[v] = fmaf(win[(G::RPI * q + w) % G::WIN][v], p.tv[0][w], a[v]);
#pragma unroll
                    for (int v = 0; v < kV; v++) a[v] *= p.gain;
                    store(oy0 + m, a);
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int UP, int DOWN, int KP, int COPY>
int launch(Params& p, const float* fx, const float* fy, int padx0, int pady0, cudaStream_t stream)
{
    typedef Geo<UP, DOWN, KP> G;
    // ---- x pass geometry and tap tables ----
    // first input column of strip 0: up 2: ceil(-padx0 / 2); else -padx0.  jRem = its residue mod 4 (the same for every strip)
    const int midS = -padx0;
    const int jFirst0 = UP == 2 ? floor_div(midS + 1, 2) : midS;
    const int jRem = pos_mod(jFirst0, 4);
    p.jA0 = jFirst0 - jRem;
    p.c0 = UP == 2 ? (jRem & 2) : 0;
    const int s = UP == 2 ? (jRem & 1) : jRem;
    for (int v = 0; v < kV; v++)
        for (int t = 0; t < kMaxNT; t++) {
            float tap = 0.f;
            if (t < G::NT) {
                const int w = G::wlo(v) + t;                   // window position
                if (UP == 2) {
                    const int mid = midS + v, a0 = pos_mod(-mid, 2);
                    const int off = (mid + a0) / 2 - jFirst0 + s;        // exact division: mid + a0 is even
                    const int k = w - off;
                    if (k >= 0 && k < KP) tap = fx[a0 + 2 * k];
                } else {
                    const int k = w - s - DOWN * v;
                    if (k >= 0 && k < G::T) tap = fx[k];
                }
            }
            p.th[v][t] = tap;
        }
    // ---- y pass ----
    for (int a = 0; a < 2; a++)
        for (int w = 0; w < kMaxWin; w++) p.tv[a][w] = 0.f;
    if (UP == 2) {
        const int mid0 = -pady0;                               // chunks start at even output rows
        if (pos_mod(mid0, 2) == 0) {
            p.iBase0 = mid0 / 2;
            for (int k = 0; k < KP; k++) { p.tv[0][k] = fy[2 * k]; p.tv[1][k + 1] = fy[1 + 2 * k]; }
        } else {
            p.iBase0 = (mid0 + 1) / 2;
            for (int k = 0; k < KP; k++) { p.tv[0][k] = fy[1 + 2 * k]; p.tv[1][k] = fy[2 * k]; }
        }
    } else {
        p.iBase0 = -pady0;
        for (int k = 0; k < G::T; k++) p.tv[0][k] = fy[k];
    }
    // ---- work split: strips of 128 columns x chunks of rows; shrink the chunks until the GPU is covered ----
    p.stripsX = (p.outW + 127) / 128;
    const long long planes = (long long)p.N * p.C;
    int chunkRows = 128;
    const long long want = (long long)sg3_sm_count() * 16 * 2;
    while (chunkRows > 32 && planes * p.stripsX * ((p.outH + chunkRows - 1) / chunkRows) < want) chunkRows >>= 1;
    p.chunkRows = chunkRows;
    p.chunksY = (p.outH + chunkRows - 1) / chunkRows;
    p.totalWarps = planes * p.stripsX * p.chunksY;
    const long long ctas = (p.totalWarps + kWarps - 1) / kWarps;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;
    const int smemBytes = kWarps * kRing * G::NINP * 4;
    auto kern = kernel<UP, DOWN, KP, COPY>;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([&] {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
    });
    if (attrErr != cudaSuccess) return (int)attrErr;
    kern<<<(unsigned)ctas, kWarps * 32, smemBytes, stream>>>(p);
    return sg3_launch_status();
}

}  // namespace ufs

// fp32 separable upfirdn2d through the streaming kernel.  fx / fy: correlation-ordered taps, zero padded to >= up * 12 entries
// (the table upfirdn2d_sep.cu builds).  SG3_E_NOKERNEL when the shape is not covered (the tiled kernel then runs).
int sg3_upfirdn2d_stream(const float* x, float* y, int N, int C, int inH, int inW, int outH, int outW,
                         const int64_t xs[4], const int64_t ys[4], const float* fx, const float* fy, int fW, int fH,
                         int up, int down, int padx0, int pady0, float gain, cudaStream_t stream)
{
    if (xs[3] != 1 || ys[3] != 1) return SG3_E_NOKERNEL;
    if (!((up == 1 || up == 2) && (down == 1 || down == 2)) || (up == 2 && down == 2)) return SG3_E_NOKERNEL;
    const int fl = fW > fH ? fW : fH;
    const int kp = (fl + up - 1) / up;
    if (kp > 12 || (up == 2 && kp > 12)) return SG3_E_NOKERNEL;
    if (((uintptr_t)x & 3) || ((uintptr_t)y & 3)) return SG3_E_NOKERNEL;
    // geometry must stay inside int32 with room for the strip / chunk arithmetic
    if ((long long)outW * 2 + 1024 > INT32_MAX || (long long)outH * 2 + 1024 > INT32_MAX) return SG3_E_NOKERNEL;
    if (padx0 < -(1 << 28) || padx0 > (1 << 28) || pady0 < -(1 << 28) || pady0 > (1 << 28)) return SG3_E_NOKERNEL;
    ufs::Params p;
    p.x = x; p.y = y; p.N = N; p.C = C; p.inH = inH; p.inW = inW; p.outH = outH; p.outW = outW;
    p.xs0 = xs[0]; p.xs1 = xs[1]; p.xs2 = xs[2]; p.ys0 = ys[0]; p.ys1 = ys[1]; p.ys2 = ys[2];
    p.gain = gain;
    auto aligned = [](const void* base, const int64_t* st, int bytes) {
        const int e = bytes / 4;
        return ((uintptr_t)base % bytes) == 0 && st[0] % e == 0 && st[1] % e == 0 && st[2] % e == 0;
    };
    const int copyBytes = aligned(x, xs, 16) ? 16 : aligned(x, xs, 8) ? 8 : 0;     // granularity of the global -> shared copies
    if (!copyBytes) return SG3_E_NOKERNEL;                                          // odd row pitch: the tiled kernel
    p.vecStore = aligned(y, ys, 16) ? 1 : 0;
#define SG3_UFS(U, D, K) return copyBytes == 16 ? ufs::launch<U, D, K, 16>(p, fx, fy, padx0, pady0, stream) \
                                                : ufs::launch<U, D, K, 8>(p, fx, fy, padx0, pady0, stream)
    if (up == 2) {
        if (kp <= 2) SG3_UFS(2, 1, 2);
        if (kp <= 4) SG3_UFS(2, 1, 4);
        if (kp <= 6) SG3_UFS(2, 1, 6);
        SG3_UFS(2, 1, 12);
    }
    if (down == 2) {
        if (kp <= 4) SG3_UFS(1, 2, 4);
        if (kp <= 8) SG3_UFS(1, 2, 8);
        SG3_UFS(1, 2, 12);
    }
    if (kp <= 4) SG3_UFS(1, 1, 4);
    if (kp <= 8) SG3_UFS(1, 1, 8);
    SG3_UFS(1, 1, 12);
#undef SG3_UFS
}
