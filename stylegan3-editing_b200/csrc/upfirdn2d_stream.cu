// upfirdn2d_stream.cu -- separable upfirdn2d (fp32; up 2, down 2 or neither; <= 12 taps per polyphase branch) as a
// warp-streaming kernel: HBM sees the input once and the output once, nothing else.
//
// The reference runs a separable filter as two launches through an intermediate image (upfirdn2d.py:241-246, kernels of
// upfirdn2d.cu:29-375); upfirdn2d_sep.cu fuses the two passes in a CTA tile but pays two __syncthreads, scalar shared-memory
// traffic and a 27 % tile halo, and ends at 0.3 of the HBM roof.  This kernel follows flrelu_stream instead:
//   * one WARP owns a strip of 128 output columns (lane = 4 adjacent columns: one 128-bit store per lane and row, a warp
//     writes 512 contiguous bytes) and streams down a chunk of rows; warps never synchronise with each other;
//   * input rows arrive in a per-warp shared-memory ring, 4-12 rows in flight per warp: by TMA (cp.async.bulk.tensor behind one
//     mbarrier per ring slot; the tensor map zero-fills everything outside the image) for the up-2 / same-rate kernels on tensors
//     TMA can address, by per-lane cp.async pieces (16 bytes when the rows allow it, else 8; out-of-image pieces written as zeros
//     by the ignore-src form) for the down-2 kernels and for rows that are not 16-byte multiples (see COPY below);
//   * the x pass reads the lane's window with 128-bit (64-bit for up 2) shared-memory loads and leaves 4 values in registers;
//   * the y pass keeps a sliding window of x-filtered rows in REGISTERS (static slots: the loop body is unrolled over one
//     rotation of the window), so the intermediate image of the reference never exists anywhere.
// All alignment cases are folded into the tap tables on the host: the staged row always starts at a column that is a multiple
// of 4 (aligned copies and loads); the 0-3 columns between that and the first column a strip needs, and the polyphase branch of
// each of a lane's 4 outputs, just shift the taps inside a slightly longer static tap range (zero taps outside the filter), and
// the taps are kernel parameters, i.e. uniform-register operands of the packed FFMA2s.  Same arithmetic as the two passes of the
// reference up to fp32 summation order (fp32 accumulation; the gain rides on the y taps).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.cuh"
#include "tensor_map.h"

namespace ufs {

// Tuning knobs (tools/build_ufs_variants.sh builds variants of this file only).  Defaults = the measured optimum on B200:
// 4 CTAs of 4 warps per SM -- more resident warps mean more concurrent row streams in DRAM and measured SLOWER (12-tap up 2:
// 0.69 of HBM at 6 CTAs / SM, 0.85 at 4) -- enforced through a lower bound on the dynamic shared memory of a CTA.
#ifndef UFS_WARPS
#define UFS_WARPS 4
#endif
#ifndef UFS_SMEM_MIN
#define UFS_SMEM_MIN 46000
#endif
#ifndef UFS_CG16
#define UFS_CG16 1
#endif
#ifndef UFS_L2PF
#define UFS_L2PF ""          // e.g. ".L2::256B": L2 prefetch size of the cp.async copies
#endif
constexpr int kWarps = UFS_WARPS;
constexpr int kV = 4;                // output columns per lane
constexpr int kMaxNT = 16, kMaxWin = 16;       // taps per output (even), y window rows

template <int UP, int DOWN, int KP> struct Geo {
    static_assert((UP == 1 || UP == 2) && (DOWN == 1 || DOWN == 2) && !(UP == 2 && DOWN == 2), "factors");
    static constexpr int T = UP * KP;                                    // taps, zero padded
    static constexpr int XSTRIDE = 128 * DOWN / UP;                      // input columns between strips
    static constexpr int LSTEP = kV * DOWN / UP;                         // input columns between lanes (2 / 8 / 4)
    // Static tap range of output v over the lane window: positions w0(v) .. w0(v) + NT - 1.  It covers the filter plus the
    // alignment slack (the 0-3 columns between the staged row's origin and the strip's first column, the polyphase branch) and
    // starts on an EVEN position with an even length, so that a packed FMA takes two adjacent window values (one 64-bit half
    // of a shared-memory load) times two adjacent taps: out = sum of the two halves of the packed accumulator.
    static constexpr int w0(int v) { return UP == 2 ? 0 : (v * DOWN) & ~1; }
    static constexpr int NT = UP == 2 ? ((KP + 3 + 1) / 2) * 2 : ((T + 3 + (DOWN == 1 ? 1 : 0) + 1) / 2) * 2;
    static constexpr int LG = UP == 2 ? 2 : 4;                           // floats per shared-memory load of the window
    static constexpr int NWH = ((w0(kV - 1) + NT + LG - 1) / LG) * LG;   // lane window (floats)
    static constexpr int NIN = ((31 * LSTEP + (UP == 2 ? 2 : 0) + NWH + 7) / 8) * 8;    // staged floats per row (even number of chunks)
    // Down 2: lanes read their window 8 floats = two 16-byte chunks apart, so a 128-bit shared-memory load of a warp would touch
    // every other chunk (bank conflicts).  The staged row is therefore DE-INTERLEAVED: even chunks first, odd chunks behind them.
    // Load j of lane L wants chunk 2 L + j, which sits at position L + j / 2 of its half: consecutive lanes read consecutive
    // 16-byte chunks, every load is one contiguous 512-byte access.  pidx() maps a float index of the row to its place; the
    // copies that fill the row use it too (pieces are at most one chunk and chunk-aligned).
    static constexpr bool SPLIT = DOWN == 2;
    static constexpr int HALFC = ((NIN / 8 + 7) / 8) * 8;               // 16-byte chunks per half, padded to 128 bytes
    static __host__ __device__ constexpr int pidx(int i)
    {
        return SPLIT ? 4 * ((i >> 3) + ((i >> 2) & 1) * HALFC) + (i & 3) : i;
    }
    static constexpr int WIN = UP == 2 ? KP + 1 : T;                     // y window (rows of x-filtered values)
    static constexpr int RPI = DOWN;                                     // input rows consumed per iteration
    static constexpr int PRIME = WIN - RPI;                              // rows consumed before the first output
    static constexpr int PERIOD = WIN / RPI;                             // iterations per rotation of the window
    // The loop body is unrolled over KU rotations of the window = ROWS_U input rows, and the ring of staged rows has RING slots
    // with RING | ROWS_U: window slot AND ring slot of every row are then compile-time constants (no index arithmetic, the
    // slot offsets are immediates of the shared-memory instructions).
    static constexpr int KU = WIN < 5 ? 2 : 1;
    static constexpr int ROWS_U = WIN * KU;
    static constexpr int RING = ROWS_U <= 8 ? ROWS_U : (ROWS_U % 6 == 0 ? 6 : ROWS_U);
    static constexpr int SLOT = SPLIT ? 8 * HALFC : ((NIN + 31) / 32) * 32;   // floats per ring slot (a multiple of 128 bytes)
    static_assert(WIN % RPI == 0 && ROWS_U % RING == 0 && NT <= kMaxNT && WIN <= kMaxWin, "window geometry");
};

struct Params {
    alignas(64) CUtensorMap mapX;   // TMA staging (COPY = 0): x as {W, H, C, N}, box {NIN, 1, 1, 1}
    const float* x; float* y;
    int N, C, inH, inW, outH, outW;
    long long xs0, xs1, xs2, ys0, ys1, ys2;      // element strides (unit stride along x)
    float gain;                                  // folded into tv by the host
    int stripsX, chunksY, chunkRows;
    long long totalWarps;
    int jA0;          // first staged input column of strip 0 (multiple of 4, may be negative)
    int c0;           // offset of lane 0's window inside the staged row (up 2: 0 or 2)
    int iBase0;       // input row at window position 0 of the chunk that starts at output row 0
    int vecStore;     // 2: output rows are 16-byte aligned (one 128-bit store per lane); 1: 8-byte aligned (two 64-bit stores); 0: scalar
    float2 th[kV][kMaxNT / 2]; // x pass: taps of output v over window positions w0(v) .. w0(v) + NT - 1, in pairs
    float tv[2][kMaxWin];      // y pass: taps over the window rows (up 2: one table per output row of a pair)
};

// 16- or 8-byte global -> shared copy; `zero`: write zeros instead and do not touch the source (the ignore-src form of cp.async).
// CG: 16-byte copies bypass L1 (.cg).  Measured on B200 (tools/build_ufs_variants.sh): .cg is worth 0.75 -> 0.91 of HBM for filter2d and
// 0.76 -> 0.85 for the 12-tap up 2, but costs the down-2 kernels 0.77 -> 0.63 (their strips re-read the row halo from L1).
template <int BYTES, bool CG>
__device__ __forceinline__ void cp_async(float* dst, const float* src, bool zero)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (BYTES == 16 && CG)
        asm volatile("{\n\t.reg .pred z;\n\tsetp.ne.b32 z, %2, 0;\n\tcp.async.cg.shared.global" UFS_L2PF " [%0], [%1], 16, z;\n\t}" ::"r"(d), "l"(src), "r"((int)zero) : "memory");
    else if (BYTES == 16)
        asm volatile("{\n\t.reg .pred z;\n\tsetp.ne.b32 z, %2, 0;\n\tcp.async.ca.shared.global" UFS_L2PF " [%0], [%1], 16, z;\n\t}" ::"r"(d), "l"(src), "r"((int)zero) : "memory");
    else
        asm volatile("{\n\t.reg .pred z;\n\tsetp.ne.b32 z, %2, 0;\n\tcp.async.ca.shared.global" UFS_L2PF " [%0], [%1], 8, z;\n\t}" ::"r"(d), "l"(src), "r"((int)zero) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// COPY: how the input rows reach the ring.  0 = TMA (cp.async.bulk.tensor behind one mbarrier per ring slot: one elected lane starts
// a row, the async proxy writes whole 128-byte lines and zero-fills everything outside the image) -- the up-2 and same-rate kernels
// on tensors TMA can address (12-tap up 2: 0.85 -> 0.88 of HBM).  16 / 8 = per-lane cp.async pieces of that many bytes: rows or
// planes that are not 16-byte multiples (e.g. 2098 columns), and every down-2 kernel -- its strips re-read a 12-column halo that
// .ca copies find in L1 while TMA goes to L2 each time (measured: downsample2d 0.76 with cp.async.ca, 0.63 with TMA).
template <int UP, int DOWN, int KP, int COPY>
__global__ void __launch_bounds__(kWarps * 32, (KP <= 6 ? 4 : 3) * 4 / kWarps) kernel(const __grid_constant__ Params p)
{
    typedef Geo<UP, DOWN, KP> G;
    extern __shared__ __align__(128) float smem[];
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
    constexpr int WARP_FLOATS = G::RING * G::SLOT + (COPY == 0 ? 2 * G::RING + 2 : 0);     // ring (+ one mbarrier per slot, 16-byte padded)
    float* ring = smem + warp * (((WARP_FLOATS + 31) / 32) * 32);

    // ---- staging: input row iBase + r -> ring slot r % RING, strictly in order ----
    int rI = 0;                                                   // next row to stage
    // cp.async path: per lane and 16- / 8-byte piece of the row, everything that does not depend on the row is computed once:
    // where the piece lands (pidx), its source column (clamped into the image so that the address is always valid and aligned) and
    // whether it lies inside the image; pieces and rows outside the image are written as zeros (the ignore-src form of cp.async).
    constexpr int CE = COPY ? COPY / 4 : 4;                       // floats per piece
    constexpr int NF = COPY ? (G::NIN / CE + 31) / 32 : 1;        // pieces per lane
    float* dstC[NF];
    unsigned srcC[NF];
    unsigned inside = 0;                                          // bit f: piece f of this lane lies inside the image (inW % CE == 0: never partly)
    bool hasLast = false;                                         // only the last piece index can be missing for some lanes
    const char* pI = nullptr;                                     // address of row rI (not dereferenced outside the image)
    const long long rowBytes = p.xs2 * 4;
    // TMA path: one mbarrier per slot; `phases` holds the parity the next wait on each slot expects
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + G::RING * G::SLOT);
    unsigned phases = 0;
    if (COPY != 0) {
#pragma unroll
        for (int f = 0; f < NF; f++) {
            const int e = CE * (lane + 32 * f), j = jA + e;
            const bool in = e < G::NIN && j >= 0 && j < p.inW;
            inside |= in ? 1u << f : 0u;
            srcC[f] = in ? 4u * (unsigned)j : 0u;
            dstC[f] = ring + (e < G::NIN ? G::pidx(e) : 0);
        }
        hasLast = CE * (lane + 32 * (NF - 1)) < G::NIN;
        pI = reinterpret_cast<const char*>(xp + (long long)iBase * p.xs2);
    } else {
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < G::RING; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[q])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    auto issue = [&](int slot) {
        if (COPY != 0) {
            if (rI < nRows) {
                const bool ok = (unsigned)(iBase + rI) < (unsigned)p.inH;
                const unsigned live = ok ? inside : 0u;
                const char* src = ok ? pI : reinterpret_cast<const char*>(xp);
#pragma unroll
                for (int f = 0; f < NF; f++)
                    if (f < NF - 1 || hasLast)
                        cp_async<COPY ? COPY : 16, UFS_CG16 != 0 && DOWN == 1>(dstC[f] + slot * G::SLOT, reinterpret_cast<const float*>(src + srcC[f]),
                                                                             !((live >> f) & 1u));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            pI += rowBytes;
        } else if (rI < nRows && lane == 0) {
            const uint32_t bar = smem_u32(&bars[slot]);
            const uint32_t dst = smem_u32(ring + slot * G::SLOT);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(G::NIN * 4)) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(dst), "l"((uint64_t)&p.mapX), "r"(bar), "r"(jA), "r"(iBase + rI), "r"(c), "r"(n) : "memory");
        }
        rI++;
    };
#pragma unroll
    for (int r = 0; r < G::RING - 1; r++) issue(r);

    // ---- x pass of the next staged row (ring slot `slot`): 4 values per lane; then refill the slot read one row ago ----
    const float* lds0 = ring + (UP == 2 ? p.c0 + G::LSTEP * lane : (G::SPLIT ? 4 * lane : G::LSTEP * lane));
    auto xrow = [&](float (&h)[kV], int slot) {
        if (COPY != 0) {
            asm volatile("cp.async.wait_group %0;" ::"n"(G::RING - 2) : "memory");
        } else {
            const uint32_t bar = smem_u32(&bars[slot]);
            for (uint32_t spins = 0; !mbar_try_wait(bar, (phases >> slot) & 1u); spins++)
                if (spins > (1u << 24)) __trap();                 // bounded: trap, never hang
            phases ^= 1u << slot;
        }
        __syncwarp();
        const float* row = lds0 + slot * G::SLOT;
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
                // split layout: chunk 2 L + w / 4 of the row sits at position L + w / 8 of its half (pidx with the lane part in lds0)
                const float4 t = *reinterpret_cast<const float4*>(row + (G::SPLIT ? G::pidx(w) : w));
                xin[w] = t.x; xin[w + 1] = t.y; xin[w + 2] = t.z; xin[w + 3] = t.w;
            }
        }
        // every lane has passed the __syncwarp above, so the slot read one row ago is free: refill it
        issue((slot + G::RING - 1) % G::RING);
#pragma unroll
        for (int v = 0; v < kV; v++) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < G::NT; t += 2)
                acc = __ffma2_rn(make_float2(xin[G::w0(v) + t], xin[G::w0(v) + t + 1]), p.th[v][t >> 1], acc);
            h[v] = acc.x + acc.y;
        }
    };

    // ---- output rows leave in order; how a lane stores is decided once ----
    //   0: nothing (column outside the image), 1: one 128-bit store, 2: two 64-bit stores (rows aligned to 8 bytes only, e.g. 2098
    //   columns), 3: scalar stores with a bound check (image edge, or rows without alignment)
    const int storeMode = ox >= p.outW ? 0 : (ox + 3 < p.outW && p.vecStore == 2) ? 1 : (ox + 3 < p.outW && p.vecStore == 1) ? 2 : 3;
    // The address of every row is computed afresh from a row counter: a running pointer updated in place right after the store
    // waits for the store to release its address register (measured: the largest single stall of the kernel).
    float* const y0 = yp + (long long)oy0 * p.ys2 + ox;
    int orow = 0;
    auto store = [&](const float2 (&o)[2]) {
        float* yq = y0 + (long long)orow * p.ys2;
        if (storeMode == 1) {
            *reinterpret_cast<float4*>(yq) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        } else if (storeMode == 2) {
            *reinterpret_cast<float2*>(yq) = o[0];
            *reinterpret_cast<float2*>(yq + 2) = o[1];
        } else if (storeMode == 3) {
            const float e[kV] = {o[0].x, o[0].y, o[1].x, o[1].y};
#pragma unroll
            for (int v = 0; v < kV; v++)
                if (ox + v < p.outW) yq[v] = e[v];
        }
        orow++;
    };

    // ---- main loop.  Iteration t consumes RPI input rows into static window slots and, once the window is full (t >= LEAD),
    //      produces the output row(s) m = t - LEAD.  The y taps carry the gain.
    constexpr int LEAD = G::PRIME / G::RPI;
    constexpr int ITS_U = G::ROWS_U / G::RPI;                    // iterations per unrolled body
    static_assert(G::PRIME % G::RPI == 0, "lead-in");
    float win[G::WIN][kV];
    const int nT = nIt + LEAD;
#pragma unroll 1
    for (int tb = 0; tb < nT; tb += ITS_U) {
#pragma unroll
        for (int q = 0; q < ITS_U; q++) {
            const int t = tb + q;
            if (t < nT) {
#pragma unroll
                for (int j = 0; j < G::RPI; j++)
                    xrow(win[(G::RPI * q + j) % G::WIN], (G::RPI * q + j) % G::RING);
                if (t >= LEAD) {
                    // y pass: packed FMAs over column pairs (v, v + 1), the tap broadcast to both halves.  Window row w of this
                    // output sits in slot (RPI q + RPI + w) % WIN (the rows just written are its last ones).
                    float2 a[2], b[2];
                    a[0] = a[1] = b[0] = b[1] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < G::WIN; w++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float* r = win[(G::RPI * q + G::RPI + w) % G::WIN];
                            const float2 hv = make_float2(r[2 * h], r[2 * h + 1]);
                            a[h] = __ffma2_rn(hv, make_float2(p.tv[0][w], p.tv[0][w]), a[h]);
                            if (UP == 2) b[h] = __ffma2_rn(hv, make_float2(p.tv[1][w], p.tv[1][w]), b[h]);
                        }
                    store(a);
                    if (UP == 2 && 2 * (t - LEAD) + 1 < chs) store(b);
                }
            }
        }
    }
    if (COPY != 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
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
                const int w = G::w0(v) + t;                    // window position
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
            if (t & 1) p.th[v][t >> 1].y = tap; else p.th[v][t >> 1].x = tap;
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
    for (int a = 0; a < 2; a++)
        for (int w = 0; w < kMaxWin; w++) p.tv[a][w] *= p.gain;           // the gain rides on the y taps
    static_assert(COPY != 0 || DOWN == 1, "TMA staging is built for the up-2 / same-rate kernels");
    if (COPY == 0) {
        // TMA staging: x as a tiled tensor map (out-of-image columns / rows read as zero).  Extent-1 dimensions get a synthetic stride.
        const uint64_t sW = 4, sH = (uint64_t)p.xs2 * 4, sC = (uint64_t)p.xs1 * 4, sN = (uint64_t)p.xs0 * 4;
        auto pad16 = [](uint64_t v) { return (v + 15) / 16 * 16; };
        const uint64_t rowS = p.inH > 1 ? sH : pad16((uint64_t)p.inW * sW);
        const uint64_t chS = p.C > 1 ? sC : pad16(rowS * (uint64_t)p.inH);
        const uint64_t smS = p.N > 1 ? sN : pad16(chS * (uint64_t)p.C);
        const uint64_t dims[4] = {(uint64_t)p.inW, (uint64_t)p.inH, (uint64_t)p.C, (uint64_t)p.N};
        const uint64_t strides[3] = {rowS, chS, smS};
        const uint32_t box[4] = {(uint32_t)G::NIN, 1, 1, 1};
        const bool ok = sg3_make_tensor_map(&p.mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (!ok) return SG3_E_NOKERNEL;
    }
    // ---- work split: strips of 128 columns x chunks of rows; shrink the chunks until the GPU is covered ----
    p.stripsX = (p.outW + 127) / 128;
    const long long planes = (long long)p.N * p.C;
    // chunks of 128 output rows; smaller ones (down to a y halo of ~10 %) until the warps fill several waves of the GPU
    int chunkRows = 128;
    const int minRows = G::T >= 8 ? 64 : 32;
    const long long want = (long long)sg3_sm_count() * 24 * 4;
    while (chunkRows > minRows && planes * p.stripsX * ((p.outH + chunkRows - 1) / chunkRows) < want) chunkRows >>= 1;
    p.chunkRows = chunkRows;
    p.chunksY = (p.outH + chunkRows - 1) / chunkRows;
    p.totalWarps = planes * p.stripsX * p.chunksY;
    const long long ctas = (p.totalWarps + kWarps - 1) / kWarps;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;
    constexpr int WARP_FLOATS = G::RING * G::SLOT + (COPY == 0 ? 2 * G::RING + 2 : 0);
    constexpr int need = kWarps * (((WARP_FLOATS + 31) / 32) * 32) * 4;
    const int smemBytes = need > UFS_SMEM_MIN ? need : UFS_SMEM_MIN;
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

template <int UP, int DOWN, int KP>
int launch_tma(Params& p, const float* fx, const float* fy, int padx0, int pady0, cudaStream_t stream)
{
    if constexpr (DOWN == 1) return launch<UP, DOWN, KP, 0>(p, fx, fy, padx0, pady0, stream);
    else return SG3_E_NOKERNEL;
}

}  // namespace ufs

// SG3_UPFIRDN_NO_TMA=1 in the environment keeps the cp.async staging for every tensor (A/B timing)
static bool ufs_tma_disabled()
{
    static const bool off = [] { const char* e = getenv("SG3_UPFIRDN_NO_TMA"); return e && e[0] == '1'; }();
    return off;
}

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
    // granularity of the global -> shared copies: rows aligned to it and made of whole pieces
    const int copyBytes = (aligned(x, xs, 16) && inW % 4 == 0) ? 16 : (aligned(x, xs, 8) && inW % 2 == 0) ? 8 : 0;
    if (!copyBytes) return SG3_E_NOKERNEL;                                          // odd row pitch: the tiled kernel
    p.vecStore = aligned(y, ys, 16) ? 2 : aligned(y, ys, 8) ? 1 : 0;
    // TMA needs 16-byte aligned rows / planes with positive pitches; everything else takes per-lane cp.async pieces
    const bool tma = down == 1 && copyBytes == 16 && xs[0] > 0 && xs[1] > 0 && xs[2] > 0 && !ufs_tma_disabled();
#define SG3_UFS(U, D, K)                                                                                                    \
    do {                                                                                                                    \
        if (tma) {                                                                                                          \
            const int rc = ufs::launch_tma<U, D, K>(p, fx, fy, padx0, pady0, stream);                                       \
            if (rc != SG3_E_NOKERNEL) return rc;                                                                            \
        }                                                                                                                   \
        return copyBytes == 16 ? ufs::launch<U, D, K, 16>(p, fx, fy, padx0, pady0, stream)                                  \
                               : ufs::launch<U, D, K, 8>(p, fx, fy, padx0, pady0, stream);                                  \
    } while (0)
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
