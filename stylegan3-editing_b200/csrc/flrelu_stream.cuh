// flrelu_stream.cuh -- fused filtered leaky-ReLU, warp-streaming kernel for sm_100a.
//
// What it computes: torch_utils/ops/filtered_lrelu.py:122-154 / filtered_lrelu.cu:139-1099 of the
// reference, for a separable up filter (UP = 2 or 4, <= 6 taps per phase) and a down-by-2 filter of
// <= 12 taps, separable or dense 12x12 (the radial filters of config R):
//   bias -> zero-insert x UP + FIR -> gain * lrelu, clamp (+ 2-bit sign codes) -> FIR + decimate by 2.
//
// How (B200-first; nothing here follows the reference's block-tile kernel):
//  * One WARP owns one strip: TW output columns x a chunk of output rows of one (n, c) plane, and
//    streams down the rows.  Warps never synchronise with each other (no __syncthreads): each has a
//    private shared-memory ring, so the ~16 resident warps per SM sit in different stages and the LSU,
//    FMA and global-load latencies of one warp hide behind the others.
//  * Per iteration a warp produces one GROUP = 4 activation rows = 2 output rows:
//      A  TMA bulk-tensor copy of the next 2 input rows into smem (fp32, aligned strides), one pair ahead;
//         otherwise global -> registers (raw bits, one iteration ahead) -> smem; zero outside the image
//      B  horizontal polyphase upsample of 2 input rows       (lane = input column)
//      C  vertical polyphase upsample + gain/lrelu/clamp/signs (lane = 2 adjacent columns)
//      D  down-by-2 FIR accumulated in registers               (lane = 2 adjacent output columns)
//    TW (58 for UP=2, 56 for UP=4) is chosen so that B and C need exactly 128 upsampled columns =
//    whole rounds of 32 lanes.
//  * All FIR arithmetic is packed FFMA2 (fma.rn.f32x2): one instruction = 2 FMAs with the tap as a
//    broadcast scalar operand.  The packed pair is always "same tap, two pixels": two input rows in B,
//    two columns in C, and in D the two activation rows (Y, Y+2) that feed output rows (o, o+1) with the
//    same filter row.  Measured on B200: 36.4 TFMA/s vs 30.3 for scalar FFMA, with half the issue slots.
//  * D never re-reads an activation: the 6 live output-row pairs per column stay in registers and retire
//    2 rows per group; instead of moving registers the tap tables rotate (uniform constant-bank offset).
//    Activations for D are laid out [column parity][column/2][4 rows permuted (0,2,1,3)] so one
//    conflict-free LDS.128 yields both row pairs of a pixel.  Dense filters that are mirror-symmetric in x
//    (the radial jinc filters) pre-add mirrored pixels: 6 taps per filter row instead of 12.
//  * The ring of horizontally-upsampled rows keeps a duplicate of its first 7 rows behind its end, so the
//    8-row window of every group is contiguous: one address register, immediate offsets.
//  * Taps travel in the launch parameters (constant bank -> uniform registers); no global filter
//    state, any stream.
//
// Roofline note: with fp32 math this op is FP32-pipe bound on B200 (>= 84 packed-pair MACs
// per output for the dense 12x12 down filter against 8 bytes of HBM traffic), so the kernel is built to
// keep the FMA pipe busy; HBM time is ~3-4x smaller than FMA time for config R.
#pragma once

#include <cuda.h>
#include <type_traits>

#include "common.cuh"

namespace flrelu_stream {

constexpr int kTapsPerPhase = 6;      // up filter taps per polyphase branch
constexpr int kDownTaps = 12;         // down filter taps (per axis)
constexpr int kWarpsPerCta = 4;
constexpr int kDup = 7;               // ring rows mirrored behind the ring end (window height - 1)

template <int UP> struct Geo {
    static constexpr int TW = UP == 2 ? 58 : 56;                     // output columns per strip (2 per lane)
    static constexpr int AW = 2 * (TW - 1) + kDownTaps;              // activation columns feeding one strip (126 / 122)
    static constexpr int BW = ((AW + UP - 1 + UP - 1) / UP) * UP;    // upsampled columns computed per strip (128 / 128)
    static constexpr int NM = BW / UP;                               // input columns producing them (64 / 32)
    static constexpr int TIW = NM + kTapsPerPhase;                   // input columns loaded (70 / 38)
    static constexpr int A_ITEMS = (TIW + 31) / 32;                  // prefetch registers per lane and row (3 / 2)
    static constexpr int RING = UP == 2 ? 8 : 10;                    // live rows of the upsampled ring
    static constexpr int XH = 64;                                    // slots per parity plane (>= AW/2 + 1)
    // TMA box width: the box must start on a 16-byte boundary (column multiple of 4), so up to 3 extra columns
    // are fetched on the left; 16-byte multiple: 76 / 44
    static constexpr int TIWP = ((TIW + 3 + 3) / 4) * 4;
    static constexpr int TMA_BUF = ((2 * TIWP * 4 + 127) / 128) * 128;   // one [2 rows][TIWP] landing buffer, 128-byte aligned
    // stage-A staging: register path = [TIW] float2; TMA path = two landing buffers + two mbarriers
    static constexpr int SIN_BYTES = 2 * TMA_BUF + 128;
    static constexpr int SB_BYTES = (RING + kDup) * BW * 4;
    static constexpr int SC_BYTES = 2 * XH * 16;
    static constexpr int SS_ROW = 144;                               // sign staging words (>= AW + 3, multiple of 4)
    static constexpr int SS_BYTES = 4 * SS_ROW;
    // per-warp shared memory; the sign staging area exists only in sign-WRITE kernels (5 instead of 4 CTAs per SM otherwise)
    static constexpr int warp_bytes(int mode) { return ((SIN_BYTES + SB_BYTES + SC_BYTES + (mode == SG3_SIGNS_WRITE ? SS_BYTES : 0) + 127) / 128) * 128; }
    static_assert(BW == 128 && AW / 2 + 1 <= XH, "strip geometry");
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
    long long totalStrips;
    float tu[4][kTapsPerPhase];        // tu[p][k]: up taps of phase p, pre-scaled by UP (horizontal pass)
    float tv[4][kTapsPerPhase];        // vertical pass: tu * gain (the activation gain rides on the taps)
    float fdx[kDownTaps];              // separable down taps (correlation order), horizontal pass; unused when dense
    // Down taps as seen by the 6 physical accumulator slots of stage D for each of the 3 rotations (g % 3):
    // slot i holds logical accumulator k = (i + 2*rot) % 6, which pairs with filter rows 2k (half 0) and 2k+1
    // (half 1).  fdr[rot][i][half][b] = FD'[2k + half][b] (dense), fdvr[rot][i][half] = fdx[2k + half] (separable).
    // Stored with the slot index fastest (padded to 8) so the 6 taps an inner loop consumes are one 128-bit and one
    // 64-bit uniform load.
    float fdr[3][2][kDownTaps][8];     // [rot][half][b][slot]
    float fdvr[3][2][8];               // [rot][half][slot]
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

__device__ __forceinline__ float2 ffma2(float2 a, float t, float2 c) { return __ffma2_rn(a, make_float2(t, t), c); }

__device__ __forceinline__ int swz(int xh) { return xh ^ ((xh >> 3) & 1); }

// Slot of activation column c in the [column] float2 row buffer the horizontal down pass reads with a lane stride of 4
// columns: columns are grouped by c % 4 (33 slots per group, odd so that the writers' column pairs hit different banks)
// and the readers of one tap find 32 consecutive slots -- no bank conflicts (plain [c] order is 8-way conflicted).
__device__ __forceinline__ int vslot(int c) { return (c & 3) * 33 + (c >> 2); }

// leaky ReLU + clamp of two values that already carry the gain; 2-bit sign codes in WRITE mode.
//   NONE : max3(v, slope*v, -clamp) then min(., clamp): one packed FMUL2, one FMNMX3 and one FMNMX per value
//   WRITE: same value, plus code = clamped ? 2 : negative ? 1 : 0
//   READ : backward pass: scale by {1, slope, 0} according to the stored code, no clamp
template <int MODE>
__device__ __forceinline__ float2 act2(float2 v, const Params& p, unsigned rc0, unsigned rc1, unsigned& wc0, unsigned& wc1)
{
    if (MODE == SG3_SIGNS_READ) {
        if (rc0 & 1u) v.x *= p.slope;
        if (rc0 & 2u) v.x = 0.f;
        if (rc1 & 1u) v.y *= p.slope;
        if (rc1 & 2u) v.y = 0.f;
        return v;
    }
    // lrelu(v) = max(v, slope * v) for 0 <= slope <= 1 (checked on the host); the lower clamp rides in the same 3-input max
    const float2 sv = __fmul2_rn(v, make_float2(p.slope, p.slope));
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
__global__ void __launch_bounds__(kWarpsPerCta * 32, UP == 2 ? 5 : 4) kernel(const __grid_constant__ Params p)
{
    typedef Geo<UP> G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // The warp index comes from a lane-0 broadcast so that the compiler treats everything derived from it (strip
    // geometry, loop counters, ring slots, tap-table offsets) as warp-uniform and keeps it on the uniform datapath.
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const long long strip = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (strip >= p.totalStrips) return;

    unsigned char* wsm = smem_raw + warp * G::warp_bytes(MODE);
    float2* sIn = (float2*)wsm;                                        // register path: [TIW] (row 2t, row 2t+1)
    uint64_t* sBar = (uint64_t*)(wsm + 2 * G::TMA_BUF);               // TMA path: one mbarrier per landing buffer
    float* sB = (float*)(wsm + G::SIN_BYTES);                          // [RING + kDup][BW]
    float4* sC = (float4*)(wsm + G::SIN_BYTES + G::SB_BYTES);          // [2][XH] rows (0,2,1,3) of one pixel
    unsigned* sS = (unsigned*)(wsm + G::SIN_BYTES + G::SB_BYTES + G::SC_BYTES);   // [SS_ROW] sign codes: one word per column, byte j = row j

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

    const char* xPlane = (const char*)p.x + n * p.xs[0] + c * p.xs[1];
    char* yPlane = (char*)p.y + n * p.ys[0] + c * p.ys[1];
    const float bias = p.b ? (float)ld_as<T>((const T*)((const char*)p.b + c * p.bs)) : 0.f;

    // ---- stage A: global -> registers (pair t = input rows 2t, 2t+1 of the strip) -----------------
    // The raw bits stay in registers until the next iteration (nothing consumes them earlier, so the
    // global-load latency is covered by a whole B/C/D round); bias and the zero border are applied when
    // they are stored to shared memory.  Column offsets are per-lane constants of the strip.
    unsigned pre[2][G::A_ITEMS];
    unsigned preValid = 0;                     // bit (row * A_ITEMS + r): pre[row][r] holds a real pixel
    int colOff[G::A_ITEMS];                    // byte offset of the lane's column, or -1 when outside the image
#pragma unroll
    for (int r = 0; r < G::A_ITEMS; r++) {
        const int jl = lane + 32 * r, j = jBase + jl;
        colOff[r] = (jl < G::TIW && j >= 0 && j < p.inW) ? (int)(j * p.xs[3]) : -1;
    }
    // TMA path: per-lane bias terms of the stage-B accumulators, hb[r][ph] = bias * sum_k tu[ph][k] * [input column
    // m + (ph > 0) + k is inside the image]  (the tensor map zero-fills outside columns and rows).
    float hb[(G::NM / 32) * UP];
    if (TMA) {
#pragma unroll
        for (int r = 0; r < G::NM / 32; r++)
#pragma unroll
            for (int ph = 0; ph < UP; ph++) {
                float acc0 = 0.f;
#pragma unroll
                for (int k = 0; k < kTapsPerPhase; k++) {
                    const int j = jBase + lane + 32 * r + (ph > 0 ? 1 : 0) + k;
                    if (j >= 0 && j < p.inW) acc0 += p.tu[ph][k];
                }
                hb[r * UP + ph] = __shfl_sync(0xffffffffu, acc0 * bias, lane);      // opaque: keep it in a register
            }
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[0])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[1])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const int dj = pos_mod(jBase, 4);      // the TMA box starts at column jBase - dj (a multiple of 4, possibly negative)
    auto tmaIssue = [&](int t) {           // one lane starts the bulk copy of input rows 2t, 2t+1 into buffer t & 1
        if (lane == 0) {
            const uint32_t bar = smem_u32(&sBar[t & 1]);
            const uint32_t dst = smem_u32(wsm + (t & 1) * G::TMA_BUF);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * G::TIWP * 4)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(dst), "l"((uint64_t)&p.mapX), "r"(bar), "r"(jBase - dj), "r"(iBase + 2 * t), "r"(c), "r"(n) : "memory");
        }
    };
    auto tmaWait = [&](int t) {            // every lane waits until buffer t & 1 holds pair t (bounded: trap, never hang)
        const uint32_t bar = smem_u32(&sBar[t & 1]);
        const uint32_t parity = (uint32_t)(t >> 1) & 1u;
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

    int pairSlot = 0, groupSlot = 0;      // ring slots of input row 2*nextPair and of group g's first window row

    // ---- stage B: horizontal upsample of the pair held in sIn -> ring rows 2t, 2t+1 ----------------
    // Rows landing in the first kDup ring slots are also written behind the ring end.
    auto stageB = [&](int t) {
        float* row0 = sB + pairSlot * G::BW;              // pairSlot == (2t) % RING, kept incrementally (always even)
        float* row1 = row0 + G::BW;
        const bool dup0 = pairSlot < kDup, dup1 = pairSlot + 1 < kDup;
        const float* tin = (const float*)(wsm + (t & 1) * G::TMA_BUF) + dj;     // TMA landing buffer [2][TIWP], strip column 0
        float rm0 = 0.f, rm1 = 0.f;                       // row-inside-image masks of the pair (TMA path: scale the bias term)
        if (TMA) {
            const int i0 = iBase + 2 * t;
            rm0 = (i0 >= 0 && i0 < p.inH) ? 1.f : 0.f;
            rm1 = (i0 + 1 >= 0 && i0 + 1 < p.inH) ? 1.f : 0.f;
        }
#pragma unroll
        for (int r = 0; r < G::NM / 32; r++) {
            const int m = lane + 32 * r;
            float2 v[kTapsPerPhase + 1];
#pragma unroll
            for (int q = 0; q <= kTapsPerPhase; q++) {
                if (TMA) v[q] = make_float2(tin[m + q], tin[G::TIWP + m + q]);
                else v[q] = sIn[m + q];
            }
            float2 acc[UP];
#pragma unroll
            for (int ph = 0; ph < UP; ph++) {
                acc[ph] = TMA ? make_float2(hb[r * UP + ph] * rm0, hb[r * UP + ph] * rm1) : make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < kTapsPerPhase; k++) acc[ph] = ffma2(v[k + (ph > 0 ? 1 : 0)], p.tu[ph][k], acc[ph]);
            }
            if (UP == 2) {
                const float2 o0 = make_float2(acc[0].x, acc[1].x), o1 = make_float2(acc[0].y, acc[1].y);
                *(float2*)(row0 + 2 * m) = o0;
                *(float2*)(row1 + 2 * m) = o1;
                if (dup0) *(float2*)(row0 + G::RING * G::BW + 2 * m) = o0;
                if (dup1) *(float2*)(row1 + G::RING * G::BW + 2 * m) = o1;
            } else {
                const float4 o0 = make_float4(acc[0].x, acc[1].x, acc[2 % UP].x, acc[3 % UP].x);
                const float4 o1 = make_float4(acc[0].y, acc[1].y, acc[2 % UP].y, acc[3 % UP].y);
                *(float4*)(row0 + 4 * m) = o0;
                *(float4*)(row1 + 4 * m) = o1;
                if (dup0) *(float4*)(row0 + G::RING * G::BW + 4 * m) = o0;
                if (dup1) *(float4*)(row1 + G::RING * G::BW + 4 * m) = o1;
            }
        }
    };

    // ---- stage C: vertical upsample + activation of group g -> sC (+ sign codes) --------------------
    // Per-lane store slots of the two columns of each round (constants of the strip); -1 = outside D's frame.
    int cSlot0[2], cSlot1[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int xd0 = 2 * (lane + 32 * r) - ex, xd1 = xd0 + 1;
        cSlot0[r] = (xd0 >= 0 && xd0 < G::AW) ? (xd0 & 1) * G::XH + swz(xd0 >> 1) : -1;
        cSlot1[r] = (xd1 >= 0 && xd1 < G::AW) ? (xd1 & 1) * G::XH + swz(xd1 >> 1) : -1;
        // identity shuffle: makes the value opaque so it stays in a register instead of being recomputed per group
        cSlot0[r] = __shfl_sync(0xffffffffu, cSlot0[r], lane);
        cSlot1[r] = __shfl_sync(0xffffffffu, cSlot1[r], lane);
    }
    const long long sPlane = (long long)plane * p.sH;
    // FD == 0 only: live output rows of the vertical down filter per activation column pair (slot k = output row 2g+1-k),
    // and the two finished rows of a group as [column] (row 2g-5, row 2g-4) pairs, in the memory sC uses for dense filters
    float2 vacc[7][2];
#pragma unroll
    for (int k = 0; k < 7; k++) vacc[k][0] = vacc[k][1] = make_float2(0.f, 0.f);
    float2* sV = (float2*)sC;
    // Sign READ: raw sign bytes of the 4 activation rows of group g (lane l: the two bytes covering pixels 4l .. 4l+7 of the
    // strip's upsampled columns, byte-aligned base), fetched one group ahead and carried across the loop so that their
    // global-load latency never sits in front of the activation.
    unsigned sLo[4] = {0u, 0u, 0u, 0u}, sHi[4] = {0u, 0u, 0u, 0u};
    auto loadSigns = [&](int g) {
        if (MODE == SG3_SIGNS_READ) {
            const int byte0 = ((Xs - ex + p.sx) >> 2) + lane;      // arithmetic shift: floor for negative coordinates
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int sY = Ys + 4 * g + j + p.sy;
                sLo[j] = sHi[j] = 0u;
                if (sY >= 0 && sY < p.sH) {
                    const uint8_t* srow = p.s + (sPlane + sY) * p.sWb;
                    if (byte0 >= 0 && byte0 < p.sWb) sLo[j] = __ldg(srow + byte0);
                    if (byte0 + 1 >= 0 && byte0 + 1 < p.sWb) sHi[j] = __ldg(srow + byte0 + 1);
                }
            }
        }
    };
    loadSigns(0);
    auto stageC = [&](int g, auto EYc) {
        constexpr int EY = decltype(EYc)::value;
        const float* win = sB + groupSlot * G::BW;       // 8 contiguous window rows (ring + mirrored tail)
        // Sign READ: per row, lane l fetches the two sign bytes that cover pixels 4l .. 4l+7 of the strip's upsampled
        // columns (byte-aligned base); any column pair's codes are then 4 bits of one lane's 16-bit window.
        unsigned signWin[4] = {0u, 0u, 0u, 0u};
        const int signX0 = Xs - ex + p.sx;               // sign-tensor x of upsampled column 0 of the strip
        const int signOff = signX0 & 3;                  // pixel offset inside the first byte (arithmetic & also for negatives)
        if (MODE == SG3_SIGNS_READ) {
            // the raw bytes were fetched at the end of the previous group (loadSigns): their global latency is long gone
#pragma unroll
            for (int j = 0; j < 4; j++) signWin[j] = sLo[j] | (sHi[j] << 8);
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int xp = 2 * (lane + 32 * r);
            float2 w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) w[q] = *(const float2*)(win + q * G::BW + xp);
            float2 v[4];
            unsigned code0[4], code1[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int yq = j + EY;                  // row offset in the UP-aligned grid
                const int ph = yq % UP;                 // compile-time after unrolling
                const int start = yq / UP + (ph > 0 ? 1 : 0);
                float2 u = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < kTapsPerPhase; k++) u = ffma2(w[start + k], p.tv[ph][k], u);
                unsigned rc0 = 0, rc1 = 0;
                if (MODE == SG3_SIGNS_READ) {           // both columns' codes out of the row's sign window (see below)
                    const int src = min((xp + signOff) >> 2, 31);      // lane 31's window also covers pixels 128..131
                    const unsigned wbits = __shfl_sync(0xffffffffu, signWin[j], src);
                    const unsigned four = wbits >> (2 * (xp + signOff - 4 * src));
                    rc0 = four & 3u;
                    rc1 = (four >> 2) & 3u;
                }
                v[j] = act2<MODE>(u, p, rc0, rc1, code0[j], code1[j]);
            }
            if (FD == 0) {
                // separable down filter: its vertical half is applied right here, from registers.  Activation row 4g + j
                // feeds output rows 2g + (j >> 1) - k through tap (j & 1) + 2k, k = 0..5 -> accumulator slot k + 1 - (j >> 1).
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        vacc[k + 1 - (j >> 1)][r] = ffma2(v[j], p.fdx[(j & 1) + 2 * k], vacc[k + 1 - (j >> 1)][r]);
            }
            if (cSlot0[r] >= 0) {
                if (FD != 0) sC[cSlot0[r]] = make_float4(v[0].x, v[2].x, v[1].x, v[3].x);
                if (MODE == SG3_SIGNS_WRITE) {
                    const int xd0 = xp - ex;
                    sS[xd0] = code0[0] | (code0[1] << 8) | (code0[2] << 16) | (code0[3] << 24);
                }
            }
            if (cSlot1[r] >= 0) {
                if (FD != 0) sC[cSlot1[r]] = make_float4(v[0].y, v[2].y, v[1].y, v[3].y);
                if (MODE == SG3_SIGNS_WRITE) {
                    const int xd1 = xp - ex + 1;
                    sS[xd1] = code1[0] | (code1[1] << 8) | (code1[2] << 16) | (code1[3] << 24);
                }
            }
        }
        if (FD == 0) {
            // rows 2g-5 (slot 6) and 2g-4 (slot 5) are complete: publish them as (row, row) pairs per column for the horizontal
            // pass of stage D, then slide the accumulators by two output rows
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int xd0 = 2 * (lane + 32 * r) - ex;
                if (xd0 >= 0 && xd0 < G::BW) sV[vslot(xd0)] = make_float2(vacc[6][r].x, vacc[5][r].x);
                if (xd0 + 1 >= 0 && xd0 + 1 < G::BW) sV[vslot(xd0 + 1)] = make_float2(vacc[6][r].y, vacc[5][r].y);
#pragma unroll
                for (int k = 6; k >= 2; k--) vacc[k][r] = vacc[k - 2][r];
                vacc[0][r] = vacc[1][r] = make_float2(0.f, 0.f);
            }
        }
        loadSigns(g + 1);
    };

    // sign bytes this strip owns: columns [0, ownW) of D's frame (whole bytes: Xs + sx is a multiple of 4),
    // rows [0, ownH) -- the last strip/chunk also owns the filter tail.
    const int ownW = (sxi == p.stripsX - 1) ? 2 * (tws - 1) + kDownTaps : 2 * G::TW;
    const int ownH = (cyi == p.chunksY - 1) ? 2 * (chs - 1) + kDownTaps : 2 * p.chunkRows;
    // lane l owns sign byte l of the strip (columns 4l .. 4l+3 of D's frame), if the strip owns that byte at all
    const int sgnByte = ((Xs + p.sx) >> 2) + lane;
    const bool sgnLane = lane < ((ownW + 3) >> 2) && sgnByte >= 0 && sgnByte < p.sWb;
    uint8_t* sgnPtr = p.s + (sPlane + Ys + p.sy) * p.sWb + sgnByte;       // row 4g + j of the strip: + (4g + j) * sWb
    auto flushSigns = [&](int g) {
        if (MODE != SG3_SIGNS_WRITE) return;
        if (sgnLane) {
            // four columns x four rows -> one word whose byte j is the packed sign byte of row j
            // (& 0x03030303: staging words of columns this strip never computes are uninitialised)
            const uint4 c4 = *(const uint4*)(sS + 4 * lane);
            const unsigned t = (c4.x & 0x03030303u) | ((c4.y & 0x03030303u) << 2) | ((c4.z & 0x03030303u) << 4) | ((c4.w & 0x03030303u) << 6);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rd = 4 * g + j;
                const int sY = Ys + rd + p.sy;
                if (rd < ownH && sY >= 0 && sY < p.sH) sgnPtr[(long long)j * p.sWb] = (uint8_t)(t >> (8 * j));
            }
        }
        sgnPtr += 4LL * p.sWb;
    };

    // ---- stage D: down-by-2 FIR, accumulated in registers -------------------------------------------
    // Physical accumulator slot i holds one pair of output rows (o, o+1) of output column 2*lane+c from the group
    // that first touches it until it retires; its logical index k = 2g - o grows by 2 per group, so instead of
    // moving registers the taps rotate: slot i uses the tap rows of logical k = (i + 2*rot) % 6, rot = g % 3,
    // fetched from rotated tables in the constant bank with a uniform offset.  One copy of the FMA code.
    float2 acc[6][2];
    float carry[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 6; k++) { acc[k][0] = make_float2(0.f, 0.f); acc[k][1] = make_float2(0.f, 0.f); }
    const int dl = min(lane, G::TW / 2 - 1);             // idle lanes read a valid column
    int dSlot[7];                                        // swizzled slots of the 7 pixel pairs a lane reads
#pragma unroll
    for (int h = 0; h < 7; h++) dSlot[h] = __shfl_sync(0xffffffffu, swz(2 * dl + h), lane);   // opaque, see cSlot

    // store address of output row 2g-5 (the first row retired by group g), column 2*lane of this strip
    char* outRow = yPlane + (long long)(oy0 - 5) * p.ys[2] + (long long)(ox0 + 2 * lane) * p.ys[3];
    float ySum = 0.f;                   // sum of the outputs this lane stored (sign-READ kernels = backward pass only)
    auto stageD = [&](int g, auto ROTc) {
        constexpr int rot = decltype(ROTc)::value;        // compile-time rotation: immediate tap-table offsets, branch-free retire
        const float4* planeE = sC;
        const float4* planeO = sC + G::XH;
        if (FD == 2) {
            // x-symmetric dense filter: per row pair, add mirrored pixels first (12 FADD2), then 6 taps per filter row.
            const float2* pE = (const float2*)planeE;
            const float2* pO = (const float2*)planeO;
#pragma unroll
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
                        for (int i = 0; i < 6; i++) acc[i][cc] = ffma2(sm, p.fdr[rot][half][b][i], acc[i][cc]);
                    }
                }
            }
        } else if (FD == 1) {
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
                            acc[i][cc] = ffma2(pa, p.fdr[rot][0][b][i], acc[i][cc]);
                            acc[i][cc] = ffma2(pb, p.fdr[rot][1][b][i], acc[i][cc]);
                        }
                    }
                }
            }
        }
        float a0, a1, b0, b1;
        if (FD == 0) {
            // horizontal half of the separable filter on the two rows stage C finished (sV): (row 2g-5, row 2g-4) per column,
            // output columns 2*lane (h0) and 2*lane + 1 (h1, two activation columns further)
            const int base = min(4 * lane, G::BW - 16);          // idle lanes read valid columns
            float2 h0 = make_float2(0.f, 0.f), h1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < kDownTaps + 2; q++) {
                const float2 v = sV[vslot(base + q)];
                if (q < kDownTaps) h0 = ffma2(v, p.fdx[q], h0);
                if (q >= 2) h1 = ffma2(v, p.fdx[q - 2], h1);
            }
            a0 = h0.x; a1 = h1.x; b0 = h0.y; b1 = h1.y;
        } else {
            // Retire output rows 2g-5 and 2g-4: logical accumulators 5 and 4 = slots (5 + 4*rot) % 6 and (4 + 4*rot) % 6;
            // the freed slots start the next group as logical 1 and 0.
#define SG3_RETIRE(S4, S5)                                                                      \
            a0 = acc[S5][0].x + carry[0]; a1 = acc[S5][1].x + carry[1];                             \
            b0 = acc[S4][0].x + acc[S5][0].y; b1 = acc[S4][1].x + acc[S5][1].y;                     \
            carry[0] = acc[S4][0].y; carry[1] = acc[S4][1].y;                                       \
            acc[S4][0] = acc[S4][1] = acc[S5][0] = acc[S5][1] = make_float2(0.f, 0.f);
            if (rot == 0) { SG3_RETIRE(4, 5) } else if (rot == 1) { SG3_RETIRE(2, 3) } else { SG3_RETIRE(0, 1) }      // resolved at compile time
#undef SG3_RETIRE
        }
        const int oA = 2 * g - 5, oB = 2 * g - 4;
        const int oxl = 2 * lane;
        if (oxl < tws) {
            const bool two = oxl + 1 < tws;
            if (oA >= 0 && oA < chs) {
                st_as<T>((T*)outRow, a0);
                if (two) st_as<T>((T*)(outRow + p.ys[3]), a1);
                if (MODE == SG3_SIGNS_READ) ySum += two ? a0 + a1 : a0;
            }
            if (oB >= 0 && oB < chs) {
                st_as<T>((T*)(outRow + p.ys[2]), b0);
                if (two) st_as<T>((T*)(outRow + p.ys[2] + p.ys[3]), b1);
                if (MODE == SG3_SIGNS_READ) ySum += two ? b0 + b1 : b0;
            }
        }
        outRow += 2 * p.ys[2];             // the lane's store address walks down two output rows per group
    };

    // ---- schedule -----------------------------------------------------------------------------------
    // Group g reads ring rows [2g, 2g+7] (UP=2) or [g, g+7] (UP=4); pairs are produced just in time.
    int nextPair = 0;
    auto producePair = [&]() {            // pair `nextPair` is in flight (TMA) or in pre[] (register path)
        if (TMA) {
            tmaWait(nextPair);
            tmaIssue(nextPair + 1);       // the other buffer was last read by the previous stage B (a __syncwarp ago)
        } else {
            storePair();
            __syncwarp();
            loadPair(nextPair + 1);       // prefetch the following pair while computing
        }
        stageB(nextPair);
        __syncwarp();
        nextPair++;
        pairSlot = pairSlot + 2 >= G::RING ? 0 : pairSlot + 2;
    };
    if (TMA) tmaIssue(0); else loadPair(0);
    // EY is a template argument of the whole loop (the polyphase row pattern of stage C is then fixed code);
    // everything else exists once.
    auto run = [&](auto EYc) {
        int rot = 0;
        for (int g = 0; g < numGroups; g++) {
            const int lastRow = (UP == 2 ? 2 * g : g) + 7;   // highest ring row group g reads
            while (2 * nextPair <= lastRow) producePair();
            stageC(g, EYc);
            __syncwarp();
            flushSigns(g);
            if (rot == 0) stageD(g, std::integral_constant<int, 0>());
            else if (rot == 1) stageD(g, std::integral_constant<int, 1>());
            else stageD(g, std::integral_constant<int, 2>());
            __syncwarp();
            rot = rot == 2 ? 0 : rot + 1;
            groupSlot += (UP == 2 ? 2 : 1);
            groupSlot -= groupSlot >= G::RING ? G::RING : 0;
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
    if (TMA) tmaWait(nextPair);           // never leave with a bulk copy still writing this warp's shared memory
    if (MODE == SG3_SIGNS_READ && p.ysum) {        // bias gradient: one fp32 atomic per strip
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ySum += __shfl_xor_sync(0xffffffffu, ySum, o);
        if (lane == 0) atomicAdd(p.ysum + c, ySum);
    }
}

}  // namespace flrelu_stream
