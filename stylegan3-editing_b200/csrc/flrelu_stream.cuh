// flrelu_stream.cuh -- fused filtered leaky-ReLU, warp-streaming kernel for sm_100a (second generation).
//
// What it computes: torch_utils/ops/filtered_lrelu.py:122-154 / filtered_lrelu.cu:139-1099 of the
// reference, for a separable up filter (UP = 2 or 4, <= 6 taps per phase) and a down-by-2 filter of
// <= 12 taps, separable or dense 12x12 (the radial filters of config R):
//   bias -> zero-insert x UP + FIR -> gain * lrelu, clamp (+ 2-bit sign codes) -> FIR + decimate by 2.
//
// How (B200-first; nothing here follows the reference's block-tile kernel):
//  * One WARP owns one strip: TW output columns x a chunk of output rows of one (n, c) plane, and
//    streams down the rows.  Warps never synchronise with each other (no __syncthreads): ~16 resident
//    warps per SM sit in different stages and hide each other's latencies.
//  * Per iteration a warp produces one GROUP = 4 activation rows = 2 output rows:
//      A  TMA bulk-tensor copies of pairs of input rows into a 4-deep shared-memory ring (fp32, aligned strides);
//         otherwise global -> registers (raw bits, one pair ahead) -> smem; zero outside the image
//      B  horizontal polyphase upsample of 2 input rows; lane = 4 ADJACENT upsampled columns
//      C  vertical polyphase upsample + gain/lrelu/clamp/signs; same lane = same 4 columns, so the rows stage B
//         produces go straight into a REGISTER window (8 / 9 rows x 4 columns) that slides down the image --
//         the horizontally upsampled rows never touch shared memory (round 1 kept them in a smem ring that
//         every group re-read four times; ncu: smem pipe 86 % busy on the separable layers, 75 % on the dense ones)
//      D  down-by-2 FIR accumulated in registers, lane = 2 adjacent output columns, fed through a small
//         double-buffered shared-memory transpose (the only smem round trip of the activation)
//    TW (58 for UP=2, 56 for UP=4) is chosen so that B and C need exactly 128 upsampled columns = 4 per lane.
//  * All FIR arithmetic is packed FFMA2 (fma.rn.f32x2): one instruction = 2 FMAs with the tap as a
//    broadcast uniform-register operand.  The packed pair is always "same tap, two pixels": two input rows in B,
//    two columns in C, and in D the two activation rows (Y, Y+2) that feed output rows (o, o+1) with the
//    same filter row.  Measured on B200: 36.4 TFMA/s vs 30.3 for scalar FFMA, with half the issue slots.
//  * D never re-reads an activation: the 6 live output-row pairs per column stay in registers and retire
//    2 rows per group; instead of moving registers the tap tables rotate (compile-time rotation, 3 copies of
//    the group body).  Dense filters that are mirror-symmetric in x (the radial jinc filters) pre-add mirrored
//    pixels: 6 taps per filter row instead of 12.  A separable down filter applies its vertical half from
//    stage C's registers (accumulator slots rotate at compile time too) and only its horizontal half in D.
//  * Sign codes are packed in registers (one 32-bit word = 4 columns x 4 rows per lane), realigned to the
//    sign tensor's byte grid with ONE shuffle, and written / read as whole bytes; no staging buffer.
//  * Taps travel in the launch parameters (constant bank -> uniform registers); no global filter
//    state, any stream.
//
// Roofline note: with fp32 math this op is FP32-pipe bound on B200 (>= 84 packed-pair MACs
// per output for the dense 12x12 down filter against 8 bytes of HBM traffic), so the kernel is built to
// keep the FMA pipe busy: everything that is not an FMA (shared-memory traffic, address arithmetic, loop
// control) is what this generation removes.
#pragma once

#include <cuda.h>
#include <type_traits>

#include "common.cuh"

namespace flrelu_stream {

constexpr int kTapsPerPhase = 6;      // up filter taps per polyphase branch
constexpr int kDownTaps = 12;         // down filter taps (per axis)
// Tuning knobs (build.py: SG3_NVCC_EXTRA / SG3_LIB_SUFFIX build a variant library).  Measured on B200, round 2: one-warp CTAs at
// 96 registers (21 warps / SM instead of 16) are 11-15 % SLOWER on the dense layers, and a rolled `half` loop in stage D
// (108 registers) 15-17 % slower: the kernel is bound by the FP32 pipe, not by occupancy.
#ifndef SG3_FL_WARPS
#define SG3_FL_WARPS 4
#endif
#ifndef SG3_FL_MINCTAS
#define SG3_FL_MINCTAS 4
#endif
#ifndef SG3_FL_HALFLOOP
#define SG3_FL_HALFLOOP 0
#endif
constexpr int kWarpsPerCta = SG3_FL_WARPS;
constexpr int kStages = 4;            // TMA landing buffers (pairs of input rows in flight) per warp
// Groups per unrolled period of the main loop.  Inside a period everything that cycles is a compile-time constant: the tap
// rotation of stage D (period 3), the transpose-buffer parity (2), the cadence of input pairs for UP = 4 (2), and the rows of
// the register window a group reads -- the window is named for a whole period and moved down once per period, not per group.
constexpr int kPeriod = 6;

template <int UP> struct Geo {
    static constexpr int TW = UP == 2 ? 58 : 56;                     // output columns per strip (2 per lane)
    static constexpr int AW = 2 * (TW - 1) + kDownTaps;              // activation columns feeding one strip (126 / 122)
    static constexpr int BW = 128;                                   // upsampled columns computed per strip: 4 per lane
    static constexpr int NM = BW / UP;                               // input columns producing them (64 / 32)
    static constexpr int TIW = NM + kTapsPerPhase;                   // input columns loaded (70 / 38)
    static constexpr int NV = UP == 2 ? 8 : 7;                       // input columns one lane reads per row in stage B
    static constexpr int A_ITEMS = (TIW + 31) / 32;                  // prefetch registers per lane and row (3 / 2)
    static constexpr int WIN = UP == 2 ? 8 : 9;                      // window rows one group may touch (8 read + the spare row of an UP = 4 pair)
    static constexpr int SHIFT = UP == 2 ? 2 : 1;                    // window rows consumed per group
    static constexpr int KEEP = UP == 2 ? 6 : 8;                     // rows carried into the next period
    // TMA box width: the box must start on a 16-byte boundary (column multiple of 4), so up to 3 extra columns
    // are fetched on the left; 16-byte multiple: 76 / 44
    static constexpr int TIWP = ((TIW + 3 + 3) / 4) * 4;
    static constexpr int TMA_BUF = ((2 * TIWP * 4 + 127) / 128) * 128;   // one [2 rows][TIWP] landing buffer, 128-byte aligned
    // stage-A staging: register path = [TIW] float2; TMA path = kStages landing buffers + kStages mbarriers
    static constexpr int SIN_BYTES = ((kStages * TMA_BUF + kStages * 8 + 127) / 128) * 128;
    // stage C -> D transpose buffer (two of them, alternating by group).  Dense filters: two parity planes of XHP
    // float4 slots, one slot = rows (0, 2, 1, 3) of one activation column; separable: 4 x VG float2 slots (two finished rows).
    // Both are indexed with a guard offset so that the columns left of D's frame (a lane's 4 columns start up to 3
    // columns early) land in unused slots instead of needing a predicate.
    static constexpr int XHP = 68;
    static constexpr int VG = 34;
    static constexpr int SC_BYTES = 2 * XHP * 16;
    static constexpr int WARP_BYTES = ((SIN_BYTES + 2 * SC_BYTES + 127) / 128) * 128;
    static_assert(AW + UP - 1 <= BW && TIW * 8 <= SIN_BYTES && 4 * VG * 8 <= SC_BYTES, "strip geometry");
    static_assert((UP == 2 ? 2 * 31 : 31) + NV <= TIW, "stage B window");
};

struct Params {
    alignas(64) CUtensorMap mapX;      // 4-D map of x {W, H, C, N}, box {TIWP, 2, 1, 1}; used by the TMA variants only
    const void* x; void* y; const void* b; uint8_t* s;
    float* ysum;                       // optional [C]: += sum of the outputs of channel c (bias gradient of the backward pass)
    int N, C, inH, inW, outH, outW;
    long long xs[4], ys[4], bs;        // byte strides
    int px0, py0;
    float gain, slope, clamp;
    int sH, sWb, sx, sy;
    int stripsX, chunksY, chunkRows;
    int vecStore;                      // bit 0: y has unit pixel stride and 8-byte (fp32) / 4-byte (fp16) aligned rows: paired stores; bit 1: round outputs to TF32
    long long totalStrips;
    SG3_TRACE_FIELD
    // tap tables are laid out for 128-bit uniform loads (LDCU.128): rows of 8 / 12 floats, 16-byte aligned
    alignas(16) float tu[4][8];        // tu[p][k]: up taps of phase p, pre-scaled by UP (horizontal pass); k < 6
    alignas(16) float tv[4][8];        // vertical pass: tu * gain (the activation gain rides on the taps)
    alignas(16) float fdx[kDownTaps];  // separable down taps (correlation order); unused when dense
    // Dense down taps as seen by the 6 physical accumulator slots of stage D for each of the 3 rotations (g % 3):
    // slot i holds logical accumulator k = (i + 2*rot) % 6, which pairs with filter rows 2k (half 0) and 2k+1
    // (half 1).  fdr[rot][half][b >> 1][(b & 1) * 6 + i] = FD'[2k + half][b]: the 12 taps of two filter columns are three
    // 128-bit uniform loads.
    alignas(16) float fdr[3][2][kDownTaps / 2][12];
};

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

// nearest TF32 value, ties away from zero (= cvt.rna.tf32.f32) as two integer operations: the FMA pipe is the busy one
__device__ __forceinline__ float round_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ float2 ffma2(float2 a, float t, float2 c) { return __ffma2_rn(a, make_float2(t, t), c); }
__device__ __forceinline__ float2 fmul2(float2 a, float t) { return __fmul2_rn(a, make_float2(t, t)); }

__device__ __forceinline__ int swz(int xh) { return xh ^ ((xh >> 3) & 1); }

// Slot of activation column xd (D's frame, >= -4) in the [column] float2 buffer the horizontal half of a separable down
// filter reads with a lane stride of 4 columns: columns are grouped by (xd + 4) % 4 (VG slots per group) so that the readers
// of one tap -- and the writers of one of a lane's 4 columns -- touch consecutive slots: no bank conflicts.
template <int VG> __device__ __forceinline__ int vslot(int xd) { return ((xd + 4) & 3) * VG + ((xd + 4) >> 2); }

// leaky ReLU + clamp of two values that already carry the gain; 2-bit sign codes in WRITE mode.
//   NONE : max3(v, slope*v, -clamp) then min(., clamp): one packed FMUL2, one FMNMX3 and one FMNMX per value
//   WRITE: same value, plus code = clamped ? 2 : negative ? 1 : 0
//   READ : backward pass: scale by {1, slope, 0} according to the stored code, no clamp
template <int MODE>
__device__ __forceinline__ float2 act2(float2 v, const Params& p, unsigned rc0, unsigned rc1, unsigned& wc0, unsigned& wc1)
{
    if (MODE == SG3_SIGNS_READ) {
        // code bit 0: times slope, bit 1: zero.  Written as predicated updates on purpose: ptxas turns them into R2P + predicated
        // FMUL / MOV (2 instructions per value); a select-the-multiplier formulation compiled to a divergent branch per value
        if (rc0 & 1u) v.x *= p.slope;
        if (rc0 & 2u) v.x = 0.f;
        if (rc1 & 1u) v.y *= p.slope;
        if (rc1 & 2u) v.y = 0.f;
        return v;
    }
    // lrelu(v) = max(v, slope * v) for 0 <= slope <= 1 (checked on the host); the lower clamp rides in the same 3-input max
    const float2 sv = fmul2(v, p.slope);
    if (MODE == SG3_SIGNS_WRITE) {
        // code = clamped ? 2 : negative ? 1 : 0.  The sign bit of v is the "negative" code (slope >= 0: lrelu keeps the
        // sign; -0.0 scales to -0.0 either way); |lrelu(v)| > clamp <=> v > clamp or slope * v < -clamp.
        const float r0 = fmaxf(v.x, sv.x), r1 = fmaxf(v.y, sv.y);
        wc0 = (fabsf(r0) > p.clamp) ? 2u : (__float_as_uint(v.x) >> 31);
        wc1 = (fabsf(r1) > p.clamp) ? 2u : (__float_as_uint(v.y) >> 31);
        return make_float2(fminf(fmaxf(r0, -p.clamp), p.clamp), fminf(fmaxf(r1, -p.clamp), p.clamp));
    }
    float r0, r1;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r0) : "f"(v.x), "f"(sv.x), "f"(-p.clamp));
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r1) : "f"(v.y), "f"(sv.y), "f"(-p.clamp));
    return make_float2(fminf(r0, p.clamp), fminf(r1, p.clamp));
}

// FD: 0 = separable down filter, 1 = dense 12x12, 2 = dense 12x12 with fd2[a][b] == fd2[a][11-b]
// (the radial filters): column pairs are pre-added, 6 taps per filter row instead of 12.
// TMA: stage A is a cp.async.bulk.tensor of the [2 rows][TIWP] input box into shared memory (zero fill outside the
// image comes from the tensor map; the bias enters as the initial value of the stage-B accumulators).  Needs fp32,
// unit pixel stride and 16-byte aligned row/plane strides; otherwise the register-prefetch path (TMA = false) runs.
template <class T, int UP, int FD, int MODE, bool TMA>
__global__ void __launch_bounds__(kWarpsPerCta * 32, TMA ? SG3_FL_MINCTAS : (3 * 4) / kWarpsPerCta) kernel(const __grid_constant__ Params p)
{
    typedef Geo<UP> G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // The warp index comes from a lane-0 broadcast so that the compiler treats everything derived from it (strip
    // geometry, loop counters, tap-table offsets) as warp-uniform and keeps it on the uniform datapath.
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const long long strip = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (strip >= p.totalStrips) return;
    if (lane == 0) SG3_TRACE_EVENT(p, 1, (unsigned long long)strip);

    unsigned char* wsm = smem_raw + warp * G::WARP_BYTES;
    float2* sIn = (float2*)wsm;                                        // register path: [TIW] (row 2t, row 2t+1)
    uint64_t* sBar = (uint64_t*)(wsm + kStages * G::TMA_BUF);         // TMA path: one mbarrier per landing buffer
    unsigned char* sCbase = wsm + G::SIN_BYTES;                        // two C -> D transpose buffers

    // ---- strip geometry --------------------------------------------------------------------------
    const int sxi = (int)(strip % p.stripsX);
    const long long rest = strip / p.stripsX;
    const int cyi = (int)(rest % p.chunksY);
    const long long plane = rest / p.chunksY;
    const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
    const int ox0 = sxi * G::TW;
    const int oy0 = cyi * p.chunkRows;
    const int tws = min(G::TW, p.outW - ox0);                // valid output columns in this strip
    const int chs = min(p.chunkRows, p.outH - oy0);          // valid output rows in this chunk
    const int Xs = 2 * ox0, Ys = 2 * oy0;                    // activation-space origin of D
    const int ex = pos_mod(Xs - p.px0, UP), ey = pos_mod(Ys - p.py0, UP);
    const int jBase = (Xs - ex - p.px0) / UP;                // exact: first input column of the strip (may be < 0)
    const int iBase = (Ys - ey - p.py0) / UP;
    const int numGroups = (2 * chs + 10 + 3) >> 2;           // activation rows 0 .. 2*(chs-1)+11 in groups of 4
    // pairs of input rows the strip consumes: UP 2: three to prime the window + one per group; UP 4: four + one per odd group
    const int numPairs = UP == 2 ? numGroups + 3 : 4 + (numGroups >> 1);

    const char* xPlane = (const char*)p.x + n * p.xs[0] + c * p.xs[1];
    char* yPlane = (char*)p.y + n * p.ys[0] + c * p.ys[1];
    const float bias = p.b ? (float)ld_as<T>((const T*)((const char*)p.b + c * p.bs)) : 0.f;

    // Lane l owns upsampled columns 4l .. 4l+3 of the strip (frame of stages B and C: column 0 is UP-aligned, D's frame
    // starts ex columns later).  UP 2: they come from input columns 2l, 2l+1 (phases 0,1,0,1); UP 4: from input column l.
    const int cbase = UP == 2 ? 2 * lane : lane;

    // ---- stage A, register path: global -> registers (pair t = input rows 2t, 2t+1 of the strip) ---
    // The raw bits stay in registers until the next pair is needed (the global-load latency is covered by a whole
    // group); bias and the zero border are applied when they are stored to shared memory.
    unsigned pre[2][G::A_ITEMS];
    unsigned preValid = 0;                     // bit (row * A_ITEMS + r): pre[row][r] holds a real pixel
    int colOff[G::A_ITEMS];                    // byte offset of the lane's column, or -1 when outside the image
#pragma unroll
    for (int r = 0; r < G::A_ITEMS; r++) {
        const int jl = lane + 32 * r, j = jBase + jl;
        colOff[r] = (jl < G::TIW && j >= 0 && j < p.inW) ? (int)(j * p.xs[3]) : -1;
    }
    // TMA path: per-lane bias terms of the stage-B accumulators, hb[c] = bias * sum_k tu[ph][k] * [input column is inside
    // the image]  (the tensor map zero-fills outside columns and rows); kept as (hb, hb) pairs = FFMA2 addends.
    float2 hb[4];
    if (TMA) {
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int ph = UP == 2 ? (cc & 1) : cc;
            const int off = UP == 2 ? (cc >> 1) + (cc & 1) : (cc > 0 ? 1 : 0);
            float acc0 = 0.f;
#pragma unroll
            for (int k = 0; k < kTapsPerPhase; k++) {
                const int j = jBase + cbase + off + k;
                if (j >= 0 && j < p.inW) acc0 += p.tu[ph][k];
            }
            const float h = __shfl_sync(0xffffffffu, acc0 * bias, lane);      // opaque: keep it in a register
            hb[cc] = make_float2(h, h);
        }
        if (lane == 0) {
#pragma unroll
            for (int sgi = 0; sgi < kStages; sgi++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[sgi])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const int dj = pos_mod(jBase, 4);      // the TMA box starts at column jBase - dj (a multiple of 4, possibly negative)
    auto tmaIssue = [&](int t) {           // one lane starts the bulk copy of input rows 2t, 2t+1 into buffer t % kStages
        if (t < numPairs && lane == 0) {
            const uint32_t bar = smem_u32(&sBar[t & (kStages - 1)]);
            const uint32_t dst = smem_u32(wsm + (t & (kStages - 1)) * G::TMA_BUF);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * G::TIWP * 4)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(dst), "l"((uint64_t)&p.mapX), "r"(bar), "r"(jBase - dj), "r"(iBase + 2 * t), "r"(c), "r"(n) : "memory");
        }
    };
    auto tmaWait = [&](int t) {            // every lane waits until buffer t % kStages holds pair t (bounded: trap, never hang)
        const uint32_t bar = smem_u32(&sBar[t & (kStages - 1)]);
        const uint32_t parity = (uint32_t)(t / kStages) & 1u;
        for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
            if (spins > (1u << 24)) __trap();
    };
    auto loadPair = [&](int t) {
        const int i0 = iBase + 2 * t;
        preValid = 0;
#pragma unroll
        for (int row = 0; row < 2; row++) {
            const int i = i0 + row;
            const bool rowOk = i >= 0 && i < p.inH;
            const char* rp = xPlane + (long long)i * p.xs[2];
#pragma unroll
            for (int r = 0; r < G::A_ITEMS; r++) {
                const bool ok = rowOk && colOff[r] >= 0;
                unsigned bits = 0;
                if (ok) {
                    if (sizeof(T) == 4) bits = __ldg((const unsigned*)(rp + colOff[r]));
                    else bits = (unsigned)__ldg((const unsigned short*)(rp + colOff[r]));
                }
                pre[row][r] = bits;
                preValid |= (ok ? 1u : 0u) << (row * G::A_ITEMS + r);
            }
        }
    };
    auto storePair = [&]() {
#pragma unroll
        for (int r = 0; r < G::A_ITEMS; r++) {
            const int jl = lane + 32 * r;
            float2 v;
#pragma unroll
            for (int row = 0; row < 2; row++) {
                float f = 0.f;
                if ((preValid >> (row * G::A_ITEMS + r)) & 1u) {
                    if (sizeof(T) == 4) f = __uint_as_float(pre[row][r]) + bias;
                    else f = __half2float(__ushort_as_half((unsigned short)pre[row][r])) + bias;
                }
                if (row == 0) v.x = f; else v.y = f;
            }
            if (jl < G::TIW) sIn[jl] = v;
        }
    };

    // ---- the register window: WIN horizontally upsampled rows x the lane's 4 columns, as two column pairs -------
    // Six copies of the group body only pay for the small separable body: measured on B200, the dense kernels lose 10-15 %
    // with a period of 6 (46 KB of hot code: instruction-cache misses), and for the register-prefetch variants ptxas stops
    // keeping the taps in uniform registers and spills kilobytes.  They use the tap-rotation period of 3.
    constexpr int PERIOD = (TMA && FD == 0 && MODE == SG3_SIGNS_NONE) ? kPeriod : 3;
    constexpr int WINP = G::WIN + G::SHIFT * (PERIOD - 1);         // rows named during one period; 8 - 10 are live at a time
    float2 w[WINP][2];
#pragma unroll
    for (int q = 0; q < WINP; q++) w[q][0] = w[q][1] = make_float2(0.f, 0.f);

    // ---- stage B: horizontal upsample of pair t -> window rows WIN-2, WIN-1 -----------------------------
    auto stageB = [&](int t, float2 (&n0)[2], float2 (&n1)[2]) {
        float2 v[G::NV];                                  // (row 2t, row 2t+1) of input columns cbase .. cbase + NV - 1
        if (TMA) {
            const float* tin = (const float*)(wsm + (t & (kStages - 1)) * G::TMA_BUF) + dj + cbase;     // landing buffer [2][TIWP]
#pragma unroll
            for (int q = 0; q < G::NV; q++) v[q] = make_float2(tin[q], tin[G::TIWP + q]);
        } else {
#pragma unroll
            for (int q = 0; q < G::NV; q++) v[q] = sIn[cbase + q];
        }
        float2 acc[4];
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int ph = UP == 2 ? (cc & 1) : cc;
            const int off = UP == 2 ? (cc >> 1) + (cc & 1) : (cc > 0 ? 1 : 0);
            acc[cc] = TMA ? hb[cc] : make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < kTapsPerPhase - 1; k++) acc[cc] = ffma2(v[off + k], p.tu[ph][k], acc[cc]);
            // last tap as two scalar FMAs: same FMA-pipe time as one packed one, but their destinations are free, so the
            // (row, row) pairs come out as the (column, column) pairs of the window without moves
            acc[cc].x = fmaf(v[off + kTapsPerPhase - 1].x, p.tu[ph][kTapsPerPhase - 1], acc[cc].x);
            acc[cc].y = fmaf(v[off + kTapsPerPhase - 1].y, p.tu[ph][kTapsPerPhase - 1], acc[cc].y);
        }
        if (TMA) {
            // rows outside the image carry no bias (they are zero padding): only the first / last pairs of a plane
            const int i0 = iBase + 2 * t;
            if (i0 < 0 || i0 + 1 >= p.inH) {
                const bool in0 = i0 >= 0 && i0 < p.inH, in1 = i0 + 1 >= 0 && i0 + 1 < p.inH;
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
                    if (!in0) acc[cc].x -= hb[cc].x;
                    if (!in1) acc[cc].y -= hb[cc].y;
                }
            }
        }
        n0[0] = make_float2(acc[0].x, acc[1].x); n0[1] = make_float2(acc[2].x, acc[3].x);
        n1[0] = make_float2(acc[0].y, acc[1].y); n1[1] = make_float2(acc[2].y, acc[3].y);
    };

    int nextPair = 0;
    auto producePair = [&](float2 (&n0)[2], float2 (&n1)[2]) {       // pair `nextPair` is in flight (TMA) or in pre[] (register path)
        if (TMA) {
            tmaIssue(nextPair + kStages - 1);     // its buffer was last read by stage B of pair nextPair - 1, a __syncwarp ago
            tmaWait(nextPair);
        } else {
            storePair();
            __syncwarp();
            loadPair(nextPair + 1);               // prefetch the following pair while computing
        }
        stageB(nextPair, n0, n1);
        if (!TMA) __syncwarp();                   // sIn is rewritten by the next storePair
        nextPair++;
    };

    // ---- stage C constants ------------------------------------------------------------------------------
    // transpose-buffer slots of the lane's 4 columns (constants of the strip): D-frame column xd = 4*lane + cc - ex
    int cSlot[4];
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
        const int xd = 4 * lane + cc - ex;
        const int s = FD == 0 ? vslot<G::VG>(xd) : (xd & 1) * G::XHP + swz((xd >> 1) + 2);
        cSlot[cc] = __shfl_sync(0xffffffffu, s, lane);   // identity shuffle: opaque, stays in a register
    }
    const long long sPlane = (long long)plane * p.sH;
    // FD == 0 only: the vertical half of the separable down filter is accumulated here, per activation column, for the 6
    // output rows in flight; physical slot of output row o = o mod 6 (compile-time inside a group: 2g mod 6 = 2 * rot).
    float2 vacc[6][2];
#pragma unroll
    for (int k = 0; k < 6; k++) vacc[k][0] = vacc[k][1] = make_float2(0.f, 0.f);

    // Sign READ: raw sign bytes of the 4 activation rows of group g (lane l: the two bytes covering pixels 4l .. 4l+7 of the
    // strip's upsampled columns, byte-aligned base), fetched one group ahead and carried across the loop so that their
    // global-load latency never sits in front of the activation.
    unsigned sLo[4] = {0u, 0u, 0u, 0u}, sHi[4] = {0u, 0u, 0u, 0u};
    const int signX0 = Xs - ex + p.sx;                   // sign-tensor x of upsampled column 0 of the strip
    const int signOff = signX0 & 3;                      // pixel offset inside the first byte (arithmetic & also for negatives)
    const int sgnRdByte = (signX0 >> 2) + lane;          // arithmetic shift: floor for negative coordinates
    const bool sgnRdLo = sgnRdByte >= 0 && sgnRdByte < p.sWb, sgnRdHi = sgnRdByte + 1 >= 0 && sgnRdByte + 1 < p.sWb;
    // row 4g + j of the strip is sgnRd + j * sWb; only dereferenced where the row and the byte exist
    const uint8_t* sgnRd = p.s + (sPlane + Ys + p.sy) * p.sWb + sgnRdByte;
    auto loadSigns = [&](int g) {
        if (MODE == SG3_SIGNS_READ) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool rowOk = (unsigned)(Ys + 4 * g + j + p.sy) < (unsigned)p.sH;
                const uint8_t* q = sgnRd + (long long)j * p.sWb;
                sLo[j] = sHi[j] = 0u;
                if (rowOk && sgnRdLo) sLo[j] = __ldg(q);
                if (rowOk && sgnRdHi) sHi[j] = __ldg(q + 1);
            }
            sgnRd += 4LL * p.sWb;
        }
    };
    loadSigns(0);

    // sign bytes this strip owns: columns [0, ownW) of D's frame (whole bytes: Xs + sx is a multiple of 4),
    // rows [0, ownH) -- the last strip/chunk also owns the filter tail.
    const int ownW = (sxi == p.stripsX - 1) ? 2 * (tws - 1) + kDownTaps : 2 * G::TW;
    const int ownH = (cyi == p.chunksY - 1) ? 2 * (chs - 1) + kDownTaps : 2 * p.chunkRows;
    // lane l owns sign byte l of the strip (columns 4l .. 4l+3 of D's frame), if the strip owns that byte at all
    const int sgnByte = ((Xs + p.sx) >> 2) + lane;
    const bool sgnLane = lane < ((ownW + 3) >> 2) && sgnByte >= 0 && sgnByte < p.sWb;
    uint8_t* sgnPtr = p.s + (sPlane + Ys + p.sy) * p.sWb + sgnByte;       // row 4g + j of the strip: + (4g + j) * sWb
    // realignment of a lane's code word (columns 4l .. 4l+3 of the B/C frame) to D's frame, which starts ex columns later:
    // byte = (own >> 2ex) | (next lane's << (8 - 2ex)), on all four row bytes of the word at once
    const unsigned sgnMaskLo = 0x01010101u * (0xffu >> (2 * ex));
    const unsigned sgnMaskHi = 0x01010101u * ((0xffu << (8 - 2 * ex)) & 0xffu);

    // ---- stage D state ------------------------------------------------------------------------------------
    // Physical accumulator slot i holds one pair of output rows (o, o+1) of output column 2*lane+c from the group
    // that first touches it until it retires; its logical index k = 2g - o grows by 2 per group, so instead of
    // moving registers the taps rotate: slot i uses the tap rows of logical k = (i + 2*rot) % 6, rot = g % 3.
    float2 acc[6][2];
    float carry[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 6; k++) { acc[k][0] = make_float2(0.f, 0.f); acc[k][1] = make_float2(0.f, 0.f); }
    const int dl = min(lane, G::TW / 2 - 1);             // idle lanes read a valid column
    int dSlot[7];                                        // swizzled slots of the 7 pixel pairs a lane reads
#pragma unroll
    for (int h = 0; h < 7; h++) dSlot[h] = __shfl_sync(0xffffffffu, swz(2 * dl + h + 2), lane);   // opaque, see cSlot
    // FD == 0: activation column vBase + q of D's frame sits in slot ((q & 3) * VG + (q >> 2) + 1) + vRow (vBase is a multiple of 4)
    const int vRow = min(lane, G::BW / 4 - 4);

    // store address of output row 2g-5 (the first row retired by group g), column 2*lane of this strip
    char* outRow = yPlane + (long long)(oy0 - 5) * p.ys[2] + (long long)(ox0 + 2 * lane) * p.ys[3];
    const int oxl = 2 * lane;
    const bool laneStores = oxl < tws, two = oxl + 1 < tws;
    float ySum = 0.f;                   // sum of the outputs this lane stored (sign-READ kernels = backward pass only)

    // ---- one group: C (vertical upsample, activation, signs, transpose) then D (down filter, retire 2 rows) -----
    auto group = [&](int g, auto EYc, auto Rc) {
        constexpr int EY = decltype(EYc)::value;
        constexpr int R = decltype(Rc)::value;            // position of the group inside the period (g % PERIOD)
        constexpr int rot = R % 3;                        // compile-time rotation: immediate tap-table offsets, branch-free retire
        constexpr int WB = R * G::SHIFT;                  // first window row of this group
        unsigned char* sCbuf = sCbase + (PERIOD % 2 == 0 ? (R & 1) : (g & 1)) * G::SC_BYTES;
        float4* sC = (float4*)sCbuf;
        float2* sV = (float2*)sCbuf;

        // ---------------- stage C ----------------
        float2 v[4][2];
        unsigned cword = 0;                               // WRITE: byte j = codes of row j, 2 bits per column of the lane
        unsigned fours[4] = {0u, 0u, 0u, 0u};             // READ: the lane's four 2-bit codes of each row
        if (MODE == SG3_SIGNS_READ) {
#pragma unroll
            for (int j = 0; j < 4; j++) fours[j] = (sLo[j] | (sHi[j] << 8)) >> (2 * signOff);
            loadSigns(g + 1);                             // a whole group ahead: the loads come from HBM
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int yq = j + EY;                        // row offset in the UP-aligned grid
            const int ph = yq % UP;                       // compile-time after unrolling
            const int start = yq / UP + (ph > 0 ? 1 : 0);
            const unsigned four = fours[j];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float2 u = fmul2(w[WB + start][h], p.tv[ph][0]);
#pragma unroll
                for (int k = 1; k < kTapsPerPhase; k++) u = ffma2(w[WB + start + k][h], p.tv[ph][k], u);
                unsigned wc0 = 0, wc1 = 0;
                v[j][h] = act2<MODE>(u, p, (four >> (4 * h)) & 3u, (four >> (4 * h + 2)) & 3u, wc0, wc1);
                if (MODE == SG3_SIGNS_WRITE) cword |= (wc0 | (wc1 << 2)) << (8 * j + 4 * h);
            }
        }
        if (FD == 0) {
            // separable down filter: its vertical half is applied right here, from registers.  Activation row 4g + j feeds
            // output rows 2g + (j >> 1) - k through tap (j & 1) + 2k, k = 0..5.  Rows 2g and 2g+1 start in this group (their
            // slots were freed by rows 2g-6 / 2g-5), rows 2g-5 and 2g-4 finish.
            float2 fin[2][2];
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const int d = (j >> 1) - k;                       // output row 2g + d
                    const int slot = (2 * rot + d + 12) % 6;
                    const float tap = p.fdx[(j & 1) + 2 * k];
                    const bool first = k == 0 && (j & 1) == 0;        // first contribution to a new row
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        vacc[slot][h] = first ? fmul2(v[j][h], tap) : ffma2(v[j][h], tap, vacc[slot][h]);
                }
                if (j == 1) { fin[0][0] = vacc[(2 * rot + 7) % 6][0]; fin[0][1] = vacc[(2 * rot + 7) % 6][1]; }     // row 2g-5
                if (j == 3) { fin[1][0] = vacc[(2 * rot + 8) % 6][0]; fin[1][1] = vacc[(2 * rot + 8) % 6][1]; }     // row 2g-4
            }
            sV[cSlot[0]] = make_float2(fin[0][0].x, fin[1][0].x);
            sV[cSlot[1]] = make_float2(fin[0][0].y, fin[1][0].y);
            sV[cSlot[2]] = make_float2(fin[0][1].x, fin[1][1].x);
            sV[cSlot[3]] = make_float2(fin[0][1].y, fin[1][1].y);
        } else {
            sC[cSlot[0]] = make_float4(v[0][0].x, v[2][0].x, v[1][0].x, v[3][0].x);
            sC[cSlot[1]] = make_float4(v[0][0].y, v[2][0].y, v[1][0].y, v[3][0].y);
            sC[cSlot[2]] = make_float4(v[0][1].x, v[2][1].x, v[1][1].x, v[3][1].x);
            sC[cSlot[3]] = make_float4(v[0][1].y, v[2][1].y, v[1][1].y, v[3][1].y);
        }
        if (MODE == SG3_SIGNS_WRITE) {
            const unsigned nxt = __shfl_down_sync(0xffffffffu, cword, 1);
            const unsigned t = ((cword >> (2 * ex)) & sgnMaskLo) | ((nxt << (8 - 2 * ex)) & sgnMaskHi);
            if (sgnLane) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int rd = 4 * g + j;
                    const int sY = Ys + rd + p.sy;
                    if (rd < ownH && sY >= 0 && sY < p.sH) sgnPtr[(long long)j * p.sWb] = (uint8_t)(t >> (8 * j));
                }
            }
            sgnPtr += 4LL * p.sWb;
        }
        __syncwarp();

        // ---------------- stage D ----------------
        float a0, a1, b0, b1;
        if (FD == 2) {
            // x-symmetric dense filter: per row pair, add mirrored pixels first (12 FADD2), then 6 taps per filter row.
            const float2* pE = (const float2*)sC;
            const float2* pO = (const float2*)(sC + G::XHP);
#if SG3_FL_HALFLOOP
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int half = 0; half < 2; half++) {          // half 0: rows (4g, 4g+2); half 1: rows (4g+1, 4g+3)
                float2 px[kDownTaps + 2];
#pragma unroll
                for (int q = 0; q < kDownTaps + 2; q++) px[q] = ((q & 1) ? pO : pE)[2 * dSlot[q >> 1] + half];
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
#pragma unroll
                    for (int b = 0; b < kDownTaps / 2; b++) {
                        const float2 sm = __fadd2_rn(px[2 * cc + b], px[2 * cc + kDownTaps - 1 - b]);
#pragma unroll
                        for (int i = 0; i < 6; i++) acc[i][cc] = ffma2(sm, p.fdr[rot][half][b >> 1][(b & 1) * 6 + i], acc[i][cc]);
                    }
                }
            }
        } else if (FD == 1) {
            const float4* planeE = sC;
            const float4* planeO = sC + G::XHP;
#pragma unroll
            for (int q = 0; q < kDownTaps + 2; q++) {         // pixel 4*lane + q of D's frame
                const float4 px = (q & 1) ? planeO[dSlot[q >> 1]] : planeE[dSlot[q >> 1]];
                const float2 pa = make_float2(px.x, px.y);  // rows 4g, 4g+2
                const float2 pb = make_float2(px.z, px.w);  // rows 4g+1, 4g+3
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int b = q - 2 * cc;               // tap column for output column 2*lane+cc
                    if (b >= 0 && b < kDownTaps) {
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            acc[i][cc] = ffma2(pa, p.fdr[rot][0][b >> 1][(b & 1) * 6 + i], acc[i][cc]);
                            acc[i][cc] = ffma2(pb, p.fdr[rot][1][b >> 1][(b & 1) * 6 + i], acc[i][cc]);
                        }
                    }
                }
            }
        }
        if (FD == 0) {
            // horizontal half of the separable filter on the two rows stage C finished (sV): (row 2g-5, row 2g-4) per column,
            // output columns 2*lane (h0) and 2*lane + 1 (h1, two activation columns further)
            float2 h0 = make_float2(0.f, 0.f), h1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < kDownTaps + 2; q++) {
                const float2 vv = sV[vRow + ((q & 3) * G::VG + (q >> 2) + 1)];
                if (q == 0) h0 = fmul2(vv, p.fdx[0]); else if (q < kDownTaps) h0 = ffma2(vv, p.fdx[q], h0);
                if (q == 2) h1 = fmul2(vv, p.fdx[0]); else if (q > 2) h1 = ffma2(vv, p.fdx[q - 2], h1);
            }
            a0 = h0.x; a1 = h1.x; b0 = h0.y; b1 = h1.y;
        } else {
            // Retire output rows 2g-5 and 2g-4: logical accumulators 5 and 4 = slots (5 + 4*rot) % 6 and (4 + 4*rot) % 6;
            // the freed slots start the next group as logical 1 and 0.
            constexpr int S5 = (5 + 4 * rot) % 6, S4 = (4 + 4 * rot) % 6;
            a0 = acc[S5][0].x + carry[0]; a1 = acc[S5][1].x + carry[1];
            b0 = acc[S4][0].x + acc[S5][0].y; b1 = acc[S4][1].x + acc[S5][1].y;
            carry[0] = acc[S4][0].y; carry[1] = acc[S4][1].y;
            acc[S4][0] = acc[S4][1] = acc[S5][0] = acc[S5][1] = make_float2(0.f, 0.f);
        }
        const int oA = 2 * g - 5;                         // rows oA and oA + 1 retire; both valid except at the chunk's ends
        if (sizeof(T) == 4 && (p.vecStore & 2)) {         // SG3_FLRELU_ROUND_TF32: the consumer is a TF32 tensor-core conv
            a0 = round_tf32(a0); a1 = round_tf32(a1); b0 = round_tf32(b0); b1 = round_tf32(b1);
        }
        if (laneStores) {
            if ((p.vecStore & 1) && two) {
                if (oA >= 0 && oA < chs) {
                    if (sizeof(T) == 4) *(float2*)outRow = make_float2(a0, a1);
                    else *(__half2*)outRow = __floats2half2_rn(a0, a1);
                    if (MODE == SG3_SIGNS_READ) ySum += a0 + a1;
                }
                if (oA + 1 >= 0 && oA + 1 < chs) {
                    if (sizeof(T) == 4) *(float2*)(outRow + p.ys[2]) = make_float2(b0, b1);
                    else *(__half2*)(outRow + p.ys[2]) = __floats2half2_rn(b0, b1);
                    if (MODE == SG3_SIGNS_READ) ySum += b0 + b1;
                }
            } else {
                if (oA >= 0 && oA < chs) {
                    st_as<T>((T*)outRow, a0);
                    if (two) st_as<T>((T*)(outRow + p.ys[3]), a1);
                    if (MODE == SG3_SIGNS_READ) ySum += two ? a0 + a1 : a0;
                }
                if (oA + 1 >= 0 && oA + 1 < chs) {
                    st_as<T>((T*)(outRow + p.ys[2]), b0);
                    if (two) st_as<T>((T*)(outRow + p.ys[2] + p.ys[3]), b1);
                    if (MODE == SG3_SIGNS_READ) ySum += two ? b0 + b1 : b0;
                }
            }
        }
        outRow += 2 * p.ys[2];             // the lane's store address walks down two output rows per group
    };

    // ---- schedule -----------------------------------------------------------------------------------
    // UP 2: group g reads upsampled rows [2g, 2g+7]: three pairs prime the window, every group adds one.
    // UP 4: group g reads rows [g, g+7]: four pairs prime it, every odd group adds one (rows g+7, g+8).
    if (TMA) {
#pragma unroll
        for (int t = 0; t < kStages - 1; t++) tmaIssue(t);
    } else {
        loadPair(0);
    }
    constexpr int PRIME = UP == 2 ? 3 : 4;
#pragma unroll
    for (int t = 0; t < PRIME; t++) {
        producePair(w[2 * t], w[2 * t + 1]);
        if (TMA) __syncwarp();
    }
    // EY is a template argument of the whole loop (the polyphase row pattern of stage C is then fixed code); the loop is
    // unrolled by one period of kPeriod groups, at the end of which the live window rows move down to the front.
    auto step = [&](int g, auto EYc, auto Rc) {
        constexpr int R = decltype(Rc)::value;
        if (UP == 2 || (PERIOD % 2 == 0 ? (R & 1) : (g & 1))) producePair(w[R * G::SHIFT + G::WIN - 2], w[R * G::SHIFT + G::WIN - 1]);
        group(g, EYc, Rc);
    };
    auto run = [&](auto EYc) {
        for (int g = 0; g < numGroups; g += PERIOD) {
            step(g, EYc, std::integral_constant<int, 0>());
            if (g + 1 >= numGroups) break;
            step(g + 1, EYc, std::integral_constant<int, 1>());
            if (g + 2 >= numGroups) break;
            step(g + 2, EYc, std::integral_constant<int, 2>());
            if (PERIOD == 6) {
                if (g + 3 >= numGroups) break;
                step(g + 3, EYc, std::integral_constant<int, (PERIOD == 6 ? 3 : 0)>());
                if (g + 4 >= numGroups) break;
                step(g + 4, EYc, std::integral_constant<int, (PERIOD == 6 ? 4 : 0)>());
                if (g + 5 >= numGroups) break;
                step(g + 5, EYc, std::integral_constant<int, (PERIOD == 6 ? 5 : 0)>());
            }
#pragma unroll
            for (int q = 0; q < G::KEEP; q++) { w[q][0] = w[q + PERIOD * G::SHIFT][0]; w[q][1] = w[q + PERIOD * G::SHIFT][1]; }
        }
    };
    if (UP == 2) {
        if (ey == 0) run(std::integral_constant<int, 0>());
        else run(std::integral_constant<int, 1>());
    } else {
        if (ey == 0) run(std::integral_constant<int, 0>());
        else if (ey == 1) run(std::integral_constant<int, 1>());
        else if (ey == 2) run(std::integral_constant<int, (UP == 4 ? 2 : 0)>());
        else run(std::integral_constant<int, (UP == 4 ? 3 : 0)>());
    }
    if (MODE == SG3_SIGNS_READ && p.ysum) {        // bias gradient: one fp32 atomic per strip
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ySum += __shfl_xor_sync(0xffffffffu, ySum, o);
        if (lane == 0) atomicAdd(p.ysum + c, ySum);
    }
    if (lane == 0) SG3_TRACE_EVENT(p, 2, (unsigned long long)strip);
}

}  // namespace flrelu_stream
