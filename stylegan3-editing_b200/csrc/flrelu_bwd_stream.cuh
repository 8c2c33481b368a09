// flrelu_bwd_stream.cuh -- fused filtered leaky-ReLU for the BACKWARD shapes of config R: a dense (non-separable)
// 12x12 up-by-2 filter followed by a separable down-by-2 (12 taps) or down-by-4 (24 taps) filter, activation either
// looked up in the forward pass's sign tensor (gradient) or computed (forward use of the same shape).
//
// This is the adjoint of the forward layers "separable up 2 / 4 -> dense radial 12x12 down 2"
// (torch_utils/ops/filtered_lrelu.py:240-269: roles of the filters swap; reference specialisations
// filtered_lrelu.cu:1261,1266-1267 "6t-upf2-downs2 / downs4").
//
// Same warp-streaming scheme as flrelu_stream.cuh (one warp = one strip, no block barriers, 4 activation rows per
// iteration, packed FFMA2 everywhere), with a different dataflow because the up filter is not separable:
//   A  global -> registers (raw, one pair of input rows ahead) -> two small rings in shared memory that hold every
//      vertically adjacent input-row pair as one float2: ring E = rows (2t, 2t+1), ring O = rows (2t+1, 2t+2)
//   U  dense polyphase upsample: each lane owns 4 adjacent upsampled columns and the 4 rows of the group; the packed
//      pair is (row j, row j+2) -- same polyphase taps, input rows one apart, i.e. one float2 of ring E or O.
//      36 taps per output, 288 FFMA2 per lane and group.  Then the sign / lrelu step in registers.
//   V  vertical half of the separable down filter, accumulated in registers straight from U's results
//      (packed pair = two adjacent columns); finished rows go to a small shared-memory row buffer
//   H  horizontal half of the down filter from that buffer (lane = output column), store.
// The upsampled image never exists in memory: per group a warp touches 2 input rows and writes 2 (down 2) or 1
// (down 4) output rows.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace flrelu_bwd_stream {

constexpr int kWarpsPerCta = 4;
constexpr int kUpTaps = 12;            // dense up filter is kUpTaps x kUpTaps, 6 x 6 per polyphase branch
constexpr int kBW = 128;               // upsampled columns computed per strip: 4 per lane

template <int DOWN> struct Geo {
    static constexpr int FDT = 6 * DOWN;                              // down filter taps (12 / 24)
    static constexpr int TW = DOWN == 2 ? 58 : 26;                    // output columns per strip
    static constexpr int AW = DOWN * (TW - 1) + FDT;                  // activation columns that feed them (126 / 124)
    static constexpr int TIW = kBW / 2 + 6;                           // input columns loaded per row (70)
    static constexpr int A_ITEMS = (TIW + 31) / 32;
    static constexpr int RING = 4;                                    // row pairs per ring
    static constexpr int ROW_BYTES = ((TIW * 8 + 15) / 16) * 16;      // one ring row: TIW float2
    static constexpr int RING_BYTES = RING * ROW_BYTES;
    static constexpr int SV_BYTES = 4 * (kBW / 4 + 1) * 8;            // finished V rows: [vslot(column)][2] floats
    static constexpr int WARP_BYTES = ((2 * RING_BYTES + SV_BYTES + 127) / 128) * 128;
    static_assert(AW + 1 <= kBW, "strip geometry");
};

struct Params {
    const void* x; void* y; const void* b; const uint8_t* s;
    float* ysum;                       // optional [C]: += sum of the outputs of channel c (the bias gradient)
    int N, C, inH, inW, outH, outW;
    long long xs[4], ys[4], bs;
    int px0, py0;
    float slope, clamp;
    int sH, sWb, sx, sy;
    int stripsX, chunksY, chunkRows;
    int vecStore;                      // y has unit pixel stride and pair-aligned rows / planes: paired stores (down 2)
    long long totalStrips;
    float tu[2][2][6][6];              // tu[py][px][ka][kb] = up^2 * gain * FU'[ay + 2ka][bx + 2kb]
    float fd[24];                      // separable down taps, correlation order, zero padded
};

__device__ __forceinline__ float2 ffma2(float2 a, float t, float2 c) { return __ffma2_rn(a, make_float2(t, t), c); }

// Slot of column c in the finished-row buffer: grouped by c % 4 so that stage H's reads (lane stride 4 columns) are
// conflict-free (see flrelu_stream.cuh, vslot).
__device__ __forceinline__ int vslot(int c) { return (c & 3) * (kBW / 4 + 1) + (c >> 2); }

template <class T, int DOWN, int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, DOWN == 2 ? 5 : 4) kernel(const __grid_constant__ Params p)
{
    typedef Geo<DOWN> G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
    const long long strip = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (strip >= p.totalStrips) return;

    unsigned char* wsm = smem_raw + warp * G::WARP_BYTES;
    unsigned char* ringE = wsm;                          // [RING][TIW] float2 (row 2t, row 2t+1)
    unsigned char* ringO = wsm + G::RING_BYTES;          // [RING][TIW] float2 (row 2t+1, row 2t+2)
    float2* sV = (float2*)(wsm + 2 * G::RING_BYTES);     // [kBW] finished V rows of this group (down 2: two rows; down 4: .x only)

    // ---- strip geometry (same conventions as flrelu_stream.cuh) ----
    const int sxi = (int)(strip % p.stripsX);
    const long long rest = strip / p.stripsX;
    const int cyi = (int)(rest % p.chunksY);
    const long long plane = rest / p.chunksY;
    const int n = (int)(plane / p.C), c = (int)(plane - (long long)n * p.C);
    const int ox0 = sxi * G::TW, oy0 = cyi * p.chunkRows;
    const int tws = min(G::TW, p.outW - ox0), chs = min(p.chunkRows, p.outH - oy0);
    const int Xs = DOWN * ox0, Ys = DOWN * oy0;
    const int ex = pos_mod(Xs - p.px0, 2), ey = pos_mod(Ys - p.py0, 2);
    const int jBase = (Xs - ex - p.px0) / 2, iBase = (Ys - ey - p.py0) / 2;
    // activation rows 0 .. DOWN*(chs-1) + FDT - 1 in groups of 4
    const int numGroups = (DOWN * (chs - 1) + G::FDT + 3) >> 2;

    const char* xPlane = (const char*)p.x + n * p.xs[0] + c * p.xs[1];
    char* yPlane = (char*)p.y + n * p.ys[0] + c * p.ys[1];
    const float bias = p.b ? (float)ld_as<T>((const T*)((const char*)p.b + c * p.bs)) : 0.f;
    const long long sPlane = (long long)plane * p.sH;

    // ---- stage A: prefetch a pair of input rows into registers, then into the E / O rings ----
    unsigned pre[2][G::A_ITEMS];
    unsigned preValid = 0;
    float prevRow[G::A_ITEMS];                 // row 2t-1 of the lane's columns (second row of the previous pair)
    int colOff[G::A_ITEMS];
#pragma unroll
    for (int r = 0; r < G::A_ITEMS; r++) {
        const int jl = lane + 32 * r, j = jBase + jl;
        colOff[r] = (jl < G::TIW && j >= 0 && j < p.inW) ? (int)(j * p.xs[3]) : -1;
        prevRow[r] = 0.f;
    }
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
    auto storePair = [&](int t) {              // pair t -> E[t % RING], and O[(t-1) % RING] = (row 2t-1, row 2t)
        float2* e = (float2*)(ringE + (t & (G::RING - 1)) * G::ROW_BYTES);
        float2* o = (float2*)(ringO + ((t + G::RING - 1) & (G::RING - 1)) * G::ROW_BYTES);
#pragma unroll
        for (int r = 0; r < G::A_ITEMS; r++) {
            const int jl = lane + 32 * r;
            float v[2];
#pragma unroll
            for (int row = 0; row < 2; row++) {
                float f = 0.f;
                if ((preValid >> (row * G::A_ITEMS + r)) & 1u) {
                    if (sizeof(T) == 4) f = __uint_as_float(pre[row][r]) + bias;
                    else f = __half2float(__ushort_as_half((unsigned short)pre[row][r])) + bias;
                }
                v[row] = f;
            }
            if (jl < G::TIW) {
                e[jl] = make_float2(v[0], v[1]);
                o[jl] = make_float2(prevRow[r], v[0]);
            }
            prevRow[r] = v[1];
        }
    };

    // ---- V accumulators: live output rows per column pair.  down 2: 7 rows (o = 2g+1 .. 2g-5), down 4: 6 rows (g .. g-5)
    constexpr int NV = DOWN == 2 ? 7 : 6;
    float2 vacc[NV][2];                        // [slot][column pair of the lane]
#pragma unroll
    for (int k = 0; k < NV; k++) vacc[k][0] = vacc[k][1] = make_float2(0.f, 0.f);

    // Sign bytes of the 4 activation rows of group g for the lane's 4 columns: the two bytes covering pixels
    // 4*lane .. 4*lane+7 from a byte-aligned base.  Loaded at the point of use they were the largest stall of the kernel
    // (global latency in every group), so the RAW bytes of the next group are fetched at the end of the current one and
    // carried across the loop; they are only combined where the activation needs them, a whole stage U later.
    unsigned sLo[4] = {0u, 0u, 0u, 0u}, sHi[4] = {0u, 0u, 0u, 0u};
    const int sgnRdByte = ((Xs - ex + p.sx) >> 2) + lane;          // arithmetic shift: floor for negative coordinates
    const bool sgnRdLo = sgnRdByte >= 0 && sgnRdByte < p.sWb, sgnRdHi = sgnRdByte + 1 >= 0 && sgnRdByte + 1 < p.sWb;
    // row 4g + j of the strip is sgnRd + j * sWb (the pointer walks down 4 rows per call); only dereferenced where row and byte exist
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

    // ---- stage U + V for group g ----
    // codes4[j]: the lane's four 2-bit codes of row j of the group, combined from the raw bytes by the CALLER before it starts the
    // next input prefetch: the byte loads and the input-row loads share a scoreboard, so a combine placed after the prefetch (where
    // the compiler put it when it lived in here) waited a full memory latency per group for loads it does not need -- one SHF
    // carried 17 % of all stall samples of the kernel (ncu source view, L10 backward).
    auto stageUV = [&](int g, auto EYc, const unsigned (&codes4)[4]) {
        constexpr int EY = decltype(EYc)::value;
        // window row r (0..7) = input row 2g + r of the strip; pair (r, r+1) lives in E (r even) or O (r odd)
        float2 acc[2][2][2];                   // [row pair jp: rows (jp, jp+2)][column block][px]
#pragma unroll
        for (int a = 0; a < 8; a++) ((float2*)acc)[a] = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const int slot = (g + (r >> 1)) & (G::RING - 1);
            const float4* rowp = (const float4*)(((r & 1) ? ringO : ringE) + slot * G::ROW_BYTES) + lane;   // columns 2*lane ..
            float2 P[8];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 t4 = rowp[q];
                P[2 * q] = make_float2(t4.x, t4.y);
                P[2 * q + 1] = make_float2(t4.z, t4.w);
            }
#pragma unroll
            for (int jp = 0; jp < 2; jp++) {
                const int yq = jp + EY;                     // row jp of the group on the 2-aligned grid
                const int py = yq & 1;
                const int start = (yq >> 1) + (py ? 1 : 0);
                const int ka = r - start;
                if (ka >= 0 && ka < 6) {
#pragma unroll
                    for (int blk = 0; blk < 2; blk++)
#pragma unroll
                        for (int px = 0; px < 2; px++)
#pragma unroll
                            for (int kb = 0; kb < 6; kb++)
                                acc[jp][blk][px] = ffma2(P[blk + px + kb], p.tu[py][px][ka][kb], acc[jp][blk][px]);
                }
            }
        }
        // activation (sign lookup or lrelu/clamp) and repack to (column, column+1) pairs per row
        float2 rowv[4][2];                                  // [row j][column pair]
        if (MODE == SG3_SIGNS_READ) loadSigns(g + 1);       // consumed a whole group from now: the bytes come from HBM
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int jp = j & 1, hi = j >> 1;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float2 a2 = acc[jp][q >> 1][q & 1];
                v[q] = hi ? a2.y : a2.x;
            }
            if (MODE == SG3_SIGNS_READ) {
                const unsigned bits = codes4[j];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned code = (bits >> (2 * q)) & 3u;
                    if (code & 1u) v[q] *= p.slope;
                    if (code & 2u) v[q] = 0.f;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float rr;                       // lrelu + lower clamp in one 3-input max (slope in [0, 1], host-checked)
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(rr) : "f"(v[q]), "f"(v[q] * p.slope), "f"(-p.clamp));
                    v[q] = fminf(rr, p.clamp);
                }
            }
            rowv[j][0] = make_float2(v[0], v[1]);
            rowv[j][1] = make_float2(v[2], v[3]);
        }
        // V: row rD = 4g + j feeds output rows through taps a = rD - DOWN*o
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                // down 2: a = (j & 1) + 2k -> o = 2g + (j >> 1) - k -> slot (k + 1 - (j >> 1));  down 4: a = j + 4k -> o = g - k -> slot k
                const int slot = DOWN == 2 ? k + 1 - (j >> 1) : k;
                const float tap = DOWN == 2 ? p.fd[(j & 1) + 2 * k] : p.fd[j + 4 * k];
                vacc[slot][0] = ffma2(rowv[j][0], tap, vacc[slot][0]);
                vacc[slot][1] = ffma2(rowv[j][1], tap, vacc[slot][1]);
            }
        }
        // retire: down 2 -> rows 2g-5 (slot 6) and 2g-4 (slot 5); down 4 -> row g-5 (slot 5); publish for stage H
        const int xd = 4 * lane - ex;                        // D-frame column of the lane's first column
        float2 outv[4];
        if (DOWN == 2) {
            outv[0] = make_float2(vacc[6][0].x, vacc[5][0].x); outv[1] = make_float2(vacc[6][0].y, vacc[5][0].y);
            outv[2] = make_float2(vacc[6][1].x, vacc[5][1].x); outv[3] = make_float2(vacc[6][1].y, vacc[5][1].y);
        } else {
            outv[0] = make_float2(vacc[5][0].x, 0.f); outv[1] = make_float2(vacc[5][0].y, 0.f);
            outv[2] = make_float2(vacc[5][1].x, 0.f); outv[3] = make_float2(vacc[5][1].y, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (xd + q >= 0 && xd + q < kBW) sV[vslot(xd + q)] = outv[q];
        // slide the accumulators
        if (DOWN == 2) {
#pragma unroll
            for (int k = NV - 1; k >= 2; k--) { vacc[k][0] = vacc[k - 2][0]; vacc[k][1] = vacc[k - 2][1]; }
            vacc[0][0] = vacc[0][1] = vacc[1][0] = vacc[1][1] = make_float2(0.f, 0.f);
        } else {
#pragma unroll
            for (int k = NV - 1; k >= 1; k--) { vacc[k][0] = vacc[k - 1][0]; vacc[k][1] = vacc[k - 1][1]; }
            vacc[0][0] = vacc[0][1] = make_float2(0.f, 0.f);
        }
    };

    // ---- stage H: horizontal down filter of the finished rows, store ----
    // store address of the first row retired by group g (down 2: row 2g-5, column 2*lane; down 4: row g-5, column lane)
    char* outRow = yPlane + (long long)(oy0 - 5) * p.ys[2] + (long long)(ox0 + (DOWN == 2 ? 2 * lane : lane)) * p.ys[3];
    float ySum = 0.f;                   // sum of the outputs this lane stored
    auto stageH = [&](int g) {
        if (DOWN == 2) {
            const int oA = 2 * g - 5, oB = 2 * g - 4;
            const int oxl = 2 * lane;
            float2 h0 = make_float2(0.f, 0.f), h1 = make_float2(0.f, 0.f);   // (row oA, row oB) of columns oxl, oxl+1
            const int base = min(4 * lane, kBW - 16);
#pragma unroll
            for (int q = 0; q < 14; q++) {
                const float2 v = sV[vslot(base + q)];
                if (q < 12) h0 = ffma2(v, p.fd[q], h0);
                if (q >= 2) h1 = ffma2(v, p.fd[q - 2], h1);
            }
            if (oxl + 1 < tws && p.vecStore) {
                if (oA >= 0 && oA < chs) {
                    if (sizeof(T) == 4) *(float2*)outRow = make_float2(h0.x, h1.x);
                    else *(__half2*)outRow = __floats2half2_rn(h0.x, h1.x);
                    ySum += h0.x + h1.x;
                }
                if (oB >= 0 && oB < chs) {
                    if (sizeof(T) == 4) *(float2*)(outRow + p.ys[2]) = make_float2(h0.y, h1.y);
                    else *(__half2*)(outRow + p.ys[2]) = __floats2half2_rn(h0.y, h1.y);
                    ySum += h0.y + h1.y;
                }
            } else if (oxl < tws) {
                const bool two = oxl + 1 < tws;
                if (oA >= 0 && oA < chs) {
                    st_as<T>((T*)outRow, h0.x);
                    if (two) st_as<T>((T*)(outRow + p.ys[3]), h1.x);
                    ySum += two ? h0.x + h1.x : h0.x;
                }
                if (oB >= 0 && oB < chs) {
                    st_as<T>((T*)(outRow + p.ys[2]), h0.y);
                    if (two) st_as<T>((T*)(outRow + p.ys[2] + p.ys[3]), h1.y);
                    ySum += two ? h0.y + h1.y : h0.y;
                }
            }
            outRow += 2 * p.ys[2];
        } else {
            const int o = g - 5;
            const int base = min(4 * lane, kBW - 24);
            float h = 0.f;
#pragma unroll
            for (int q = 0; q < 24; q++) h = fmaf(sV[vslot(base + q)].x, p.fd[q], h);
            if (lane < tws && o >= 0 && o < chs) { st_as<T>((T*)outRow, h); ySum += h; }
            outRow += p.ys[2];
        }
    };

    // ---- schedule: group g reads input rows 2g .. 2g+7 = pairs g .. g+3 (O needs pair g+3 stored as well) ----
    int nextPair = 0;
    auto producePair = [&]() {
        storePair(nextPair);
        loadPair(nextPair + 1);
        nextPair++;
    };
    loadPair(0);
    auto run = [&](auto EYc) {
        for (int g = 0; g < numGroups; g++) {
            __syncwarp();                                   // previous stage H / U are done with the ring slot about to be overwritten
            unsigned codes4[4] = {0u, 0u, 0u, 0u};
            if (MODE == SG3_SIGNS_READ) {                   // bytes loaded during the previous group (see stageUV)
#pragma unroll
                for (int j = 0; j < 4; j++) codes4[j] = (sLo[j] | (sHi[j] << 8)) >> (2 * ((Xs - ex + p.sx) & 3));
            }
            while (nextPair <= g + 3) producePair();
            __syncwarp();
            stageUV(g, EYc, codes4);
            __syncwarp();
            stageH(g);
        }
    };
    if (ey == 0) run(std::integral_constant<int, 0>());
    else run(std::integral_constant<int, 1>());
    if (p.ysum) {                                   // bias gradient: one fp32 atomic per strip
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ySum += __shfl_xor_sync(0xffffffffu, ySum, o);
        if (lane == 0) atomicAdd(p.ysum + c, ySum);
    }
}

}  // namespace flrelu_bwd_stream
