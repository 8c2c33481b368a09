// modconv_tc3.cu -- TF32 tcgen05/TMEM implicit GEMM for the 3x3 modulated conv (StyleGAN3 config T; the grouped cuDNN
// convolution of networks_stylegan3.py:59-62 with conv_kernel = 3, padding = 2; padding = 0 serves dgrad-style use).
//
//   y[n][o][oy][ox] = sum_i sum_ky sum_kx  Wn[ky][kx][o][i] * x[n][i][oy + ky - pad][ox + kx - pad]
//
// Activations are NCHW fp32, so the pixel axis is the contiguous one and a one-pixel shift of an operand is neither a
// legal TMA box start (16-byte granularity) nor a legal UMMA descriptor start.  The shifts are therefore taken on the
// ACCUMULATOR side:
//   D[M = 128 out-channels][N = NPX pixels] += A[M][K] * B[K][N]
//   A = Wn[tap] : K-major  [128 o][32 i]  SWIZZLE_128B, one 16 KB TMA box per (tap, 32-channel chunk)
//   B = x row   : N-major  [32 i][NPX px] SWIZZLE_128B_ATOM_32B, NPX/32 TMA boxes of [32 i][32 px]; the box start is
//                 4-pixel aligned, rows/columns outside the image are zero-filled by TMA (that IS the conv padding)
//   D in TMEM   : lane = out-channel, column = pixel.  The product of input column j with tap kx belongs to output column
//                 j + 2 - kx, so the MMA of tap kx simply accumulates at a column offset of the TMEM address.
//                 Measured (tools/tc3_probe.cu): the offset must be EVEN, so kx = 0 and kx = 2 share one accumulator
//                 (offsets 2 and 0) and kx = 1 gets a second one; the epilogue adds them, one register apart.
//   ky          : input row oy + ky - pad; a tile holds R output rows and R + 2 input rows per channel chunk in smem,
//                 every input row feeds up to three output-row accumulators.
// One X row group therefore serves 9 * R MMAs of K = 32, and the weights of a (tap, chunk) are used for R * (NPX - 4)
// output pixels.  Epilogue: tcgen05.ld (thread = channel) -> even/odd add -> 32x32 transpose through padded smem ->
// stores with lanes along x (128-byte lines).
//
// Warp roles (192 threads): warps 0-3 epilogue, warp 4 TMA producer (two rings: X row groups, W taps), warp 5 TMEM
// allocator + MMA issuer.  Persistent over tiles.  The kernel can double-buffer the accumulator (accStages = 2) when two
// stages fit in 512 TMEM columns, but the planner below never asks for it: one large single-buffered tile measured faster.
// Every mbarrier wait is bounded (trap instead of hang).
//
// HALF (fp16 layers, networks_stylegan3.py:61 with x.dtype == float16): fp16 activations / weights / output, kind::f16 MMAs,
// fp32 accumulation.  Same bytes everywhere: a chunk is 64 input channels of 2 bytes, an X box is [64 i][64 px] in the standard
// SWIZZLE_128B MN-major form (8-channel atoms 1024 B apart, 64-pixel blocks one box apart), K = 16 per instruction.  Box starts
// are 16-byte = 8-pixel aligned, so NPX is a multiple of 64, a tile keeps NPX - 8 columns and the padding-2 offset is 8.
#include <cuda.h>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "tensor_map.h"
#include "tc_common.cuh"


namespace {

constexpr int BK3 = 32;                 // input channels per chunk (fp32; fp16: 64 -- 128-byte K rows either way)
constexpr int kMaxWSlots = 36;          // W ring slots; one slot = one tap of one chunk: [wRows <= 128 o][32 i] fp32
constexpr int kThreads3 = 192;
constexpr int STAGE_PITCH = 36;         // transpose staging: [32 channels][36] (16-byte rows: float4 writes, conflict-free)
constexpr int kAccPitchAlign = 32;      // accumulator pitch granularity in TMEM columns (tcgen05.ld.x32 start columns)

struct Tc3Params {
    void* y;                   // float, or __half (HALF)
    int N, I, O, H, W, OH, OW, pad;
    int yPitch;                // floats between output rows (>= OW): a 16-byte multiple lets the stencil that follows use TMA
    int NPX, R, CW;            // pixels per MMA (multiple of 32), output rows per tile, accumulator pitch (>= NPX + 4)
    int S;                     // valid output columns per tile = NPX - 4
    int xoff, colBase;         // box start = tileX * S - xoff;  output ox = tileX * S + (col - colBase)
    int accStages, accStageCols, tmemCols;
    int tilesX, tilesY, tilesO, kChunks;
    int wSlots, xGroupBytes;
    int wRows, wSlotBytes;     // rows of a W slot (min(128, ceil8(O))) and its size
    int resident;              // the ring holds every (chunk, tap) of the layer: W is reloaded only when (n, o-tile) changes
    int kxStack;               // O == 32: the three kx taps of a ky are stacked along M (rows kx * 32 + o of one M = 128 MMA): 12 instead
                               // of 36 MMAs per row and chunk; the kx shifts move to the epilogue; one accumulator per row, two stages
                               // side by side in the TMEM columns.  wLoadBytes = bytes TMA delivers per W slot (96 rows; slots are 128).
    int wLoadBytes;
    int m64;                   // O <= 64: M = 64 MMAs; their accumulator occupies lanes 32q .. 32q+15 of TMEM (measured,
                               // tools/tc_m64_probe.cu), a second one fits at lane offset 16 in the SAME columns -> two stages
    long long totalTiles;
};

template <bool HALF>
__global__ void __launch_bounds__(kThreads3, 1)
modconv_tc3_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const Tc3Params p)
{
    constexpr int BKC = HALF ? 64 : BK3;               // input channels per chunk
    constexpr int BOXPX = HALF ? 64 : 32;              // pixels per X box (one 128-byte swizzle row)
    constexpr uint32_t BOXBYTES = BKC * 128u;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barWFull[kMaxWSlots], barWEmpty[kMaxWSlots], barXFull[2], barXEmpty[2], barAccFull[2], barAccEmpty[2];
    __shared__ uint32_t tmemBase;
    // warp index through a shuffle: the compiler then knows the role dispatch is warp-uniform and keeps the producer / MMA
    // loop state in uniform registers (no R2UR traffic in front of every UTMALDG / UTCHMMA)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t xRing = base;                                        // 2 groups of (R + 2) rows x NPX px x 32 ch
    const uint32_t wRing = base + 2u * (uint32_t)p.xGroupBytes;        // wSlots x 16 KB (1024-aligned: xGroupBytes % 4096 == 0)
    float* stage = reinterpret_cast<float*>(smem + (wRing + (uint32_t)(p.wSlots * p.wSlotBytes) - smem_u32(smem)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxWSlots; s++) { mbar_init(smem_u32(&barWFull[s]), 1); mbar_init(smem_u32(&barWEmpty[s]), 1); }
        for (int a = 0; a < 2; a++) {
            mbar_init(smem_u32(&barXFull[a]), 1); mbar_init(smem_u32(&barXEmpty[a]), 1);
            mbar_init(smem_u32(&barAccFull[a]), 1); mbar_init(smem_u32(&barAccEmpty[a]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"((uint32_t)p.tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmemBase, 0);

    // tile -> (sample, row block, column block, channel block); channel block fastest: the CTAs running together share x in L2
    auto decode = [&](long long t, int& n, int& oy0, int& tx, int& o0) {
        const int to = (int)(t % p.tilesO); t /= p.tilesO;
        tx = (int)(t % p.tilesX); t /= p.tilesX;
        const int ty = (int)(t % p.tilesY);
        n = (int)(t / p.tilesY);
        oy0 = ty * p.R;
        o0 = to * 128;
    };
    const int boxesPerRow = p.NPX / BOXPX;
    const uint32_t rowBytes = (uint32_t)p.NPX * 128u;                  // [32 ch][NPX px] of one input row

    if (warp == 4) {
        // ---------------- TMA producer: the whole warp runs the loop, one elected lane issues ----------------
        uint32_t xIt = 0, wIt = 0;
        int wN = -1, wO = -1;                                     // (sample, channel block) whose weights sit in a resident ring
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x) {
            int n, oy0, tx, o0;
            decode(t, n, oy0, tx, o0);
            const int xs = tx * p.S - p.xoff;
            const bool reuseW = p.resident && n == wN && o0 == wO;
            wN = n; wO = o0;
            for (int c = 0; c < p.kChunks; c++, xIt++) {
                const uint32_t xg = xIt & 1;
                if (xIt >= 2) mbar_wait(smem_u32(&barXEmpty[xg]), ((xIt >> 1) - 1) & 1);
                const uint32_t xfull = smem_u32(&barXFull[xg]);
                mbar_expect_tx_elect(xfull, (uint32_t)p.xGroupBytes);
                const uint32_t xDst = xRing + xg * (uint32_t)p.xGroupBytes;
                for (int rr = 0; rr < p.R + 2; rr++)
                    for (int j = 0; j < boxesPerRow; j++)
                        tma_load_4d_elect(xDst + rr * rowBytes + j * BOXBYTES, &mapX, xfull, xs + BOXPX * j, oy0 - p.pad + rr, c * BKC, n);
                const int nTapLoads = p.kxStack ? 3 : 9;                  // kxStack: one box of [3 taps][32 o] rows per ky
                for (int tap = 0; tap < nTapLoads; tap++, wIt++) {
                    const uint32_t ws = wIt % (uint32_t)p.wSlots, round = wIt / (uint32_t)p.wSlots;
                    if (round > 0) mbar_wait(smem_u32(&barWEmpty[ws]), (round - 1) & 1);
                    const uint32_t wfull = smem_u32(&barWFull[ws]);
                    mbar_expect_tx_elect(wfull, reuseW ? 0u : (uint32_t)p.wLoadBytes);    // resident: hand the slot over as is
                    if (!reuseW) tma_load_4d_elect(wRing + ws * (uint32_t)p.wSlotBytes, &mapW, wfull, c * BKC, o0, p.kxStack ? 3 * tap : tap, n);
                }
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues (tc_common.cuh) ----------------
        // D = F32, A = B = TF32, A K-major (bit 15 = 0), B MN-major (bit 16 = 1), N = NPX, M = 128
        // (HALF: A = B = F16 -> format fields 0; B in the standard SWIZZLE_128B MN-major form: LBO = one box, SBO = one 8-channel atom)
        const uint32_t idesc = (1u << 4) | (HALF ? 0u : (2u << 7) | (2u << 10)) | (0u << 15) | (1u << 16) |
                               ((uint32_t)(p.NPX >> 3) << 17) | ((uint32_t)((p.m64 ? 64 : 128) >> 4) << 24);
        const uint64_t dA = umma_desc(wRing, 16, 1024);
        const uint64_t dB = HALF ? umma_desc(xRing, BOXBYTES, 1024) : umma_desc(xRing, BK3 * 128, 512, kLayoutSw128Base32);
        const uint32_t aLo0 = (uint32_t)dA, aHi = (uint32_t)(dA >> 32), bLo0 = (uint32_t)dB, bHi = (uint32_t)(dB >> 32);
        const uint32_t rowStep = rowBytes >> 4, groupStep = (uint32_t)p.xGroupBytes >> 4;
        uint32_t xIt = 0, wIt = 0, tc = 0;
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x, tc++) {
            const uint32_t as = p.accStages == 2 ? (tc & 1) : 0, use = p.accStages == 2 ? (tc >> 1) : tc;
            if (use > 0) mbar_wait(smem_u32(&barAccEmpty[as]), (use - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = p.m64 ? tmem + (as ? (16u << 16) : 0u) : tmem + as * (uint32_t)p.accStageCols;
            for (int c = 0; c < p.kChunks; c++, xIt++) {
                const uint32_t xg = xIt & 1;
                mbar_wait(smem_u32(&barXFull[xg]), (xIt >> 1) & 1);
                const uint32_t bLoG = bLo0 + xg * groupStep;
                uint32_t ws = wIt % (uint32_t)p.wSlots, wPhase = (wIt / (uint32_t)p.wSlots) & 1;
                if (p.kxStack) {
                    // A = [3 kx][32 o] rows of tap row ky (M = 128, the last 32 rows are never read back), one accumulator per output row
#pragma unroll 1
                    for (int ky = 0; ky < 3; ky++) {
                        mbar_wait(smem_u32(&barWFull[ws]), wPhase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t aLo = aLo0 + ws * ((uint32_t)p.wSlotBytes >> 4);
                        const uint32_t first = (c == 0 && ky == 0) ? 0u : 1u;
                        for (int oyl = 0; oyl < p.R; oyl++) {
                            if (HALF) umma_f16_x4<2, 128>(acc + (uint32_t)(oyl * p.CW), aLo, aHi, bLoG + (uint32_t)(oyl + ky) * rowStep, bHi, idesc, first);
                            else umma_tf32_x4<2, 64>(acc + (uint32_t)(oyl * p.CW), aLo, aHi, bLoG + (uint32_t)(oyl + ky) * rowStep, bHi, idesc, first);
                        }
                        umma_commit_elect(smem_u32(&barWEmpty[ws]));
                        if (++ws == (uint32_t)p.wSlots) { ws = 0; wPhase ^= 1; }
                    }
                    wIt += 3;
                    umma_commit_elect(smem_u32(&barXEmpty[xg]));
                    continue;
                }
#pragma unroll 1
                for (int ky = 0; ky < 3; ky++) {
#pragma unroll
                    for (int kx = 0; kx < 3; kx++) {
                        mbar_wait(smem_u32(&barWFull[ws]), wPhase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t aLo = aLo0 + ws * ((uint32_t)p.wSlotBytes >> 4);
                        const uint32_t first = (c == 0 && ky == 0 && kx < 2) ? 0u : 1u;     // first touch of an accumulator
                        const uint32_t dOff = (uint32_t)((kx & 1) * p.CW + (kx == 0 ? 2 : 0));
                        for (int oyl = 0; oyl < p.R; oyl++) {
                            if (HALF) umma_f16_x4<2, 128>(acc + dOff + (uint32_t)(2 * oyl * p.CW), aLo, aHi,
                                                          bLoG + (uint32_t)(oyl + ky) * rowStep, bHi, idesc, first);
                            else umma_tf32_x4<2, 64>(acc + dOff + (uint32_t)(2 * oyl * p.CW), aLo, aHi,
                                                     bLoG + (uint32_t)(oyl + ky) * rowStep, bHi, idesc, first);
                        }
                        umma_commit_elect(smem_u32(&barWEmpty[ws]));
                        if (++ws == (uint32_t)p.wSlots) { ws = 0; wPhase ^= 1; }
                    }
                }
                wIt += 9;
                umma_commit_elect(smem_u32(&barXEmpty[xg]));
            }
            umma_commit_elect(smem_u32(&barAccFull[as]));
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> transpose in smem -> global (warps 0-3) ----------------
        // A warp can only read its own 32 TMEM lanes (= 32 channels), but once a 32-column block is staged in shared memory
        // any warp can store it: the four warps share the channel rows of every block, so narrow layers (O = 32: one
        // warp owns all the channels) still store with 128 threads.
        uint32_t tc = 0;
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x, tc++) {
            int n, oy0, tx, o0;
            decode(t, n, oy0, tx, o0);
            const uint32_t as = p.accStages == 2 ? (tc & 1) : 0, use = p.accStages == 2 ? (tc >> 1) : tc;
            mbar_wait(smem_u32(&barAccFull[as]), use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + ((uint32_t)(32 * warp) << 16) + (p.m64 ? 0u : as * (uint32_t)p.accStageCols);
            if (p.kxStack) {
                // Warp kx (0..2) holds D_kx[o = lane][j]; output column c = sum_kx D_kx[c - 2 + kx].  Each warp stages its block already
                // shifted to output columns (the 2 - kx values that come from the previous block travel in registers), the store phase
                // adds the three planes.  Plane kx = rows 32 kx .. 32 kx + 31 of the staging buffer.
                typedef typename std::conditional<HALF, __half, float>::type OutT;
                const int colEnd = p.colBase + p.S;
                const size_t chStep = (size_t)p.OH * p.yPitch;
                float* st = stage + (32 * warp + lane) * STAGE_PITCH;
                for (int oyl = 0; oyl < p.R; oyl++) {
                    const int oy = oy0 + oyl;
                    if (oy >= p.OH) break;
                    float prev0 = 0.f, prev1 = 0.f;                         // D[c0 - 2], D[c0 - 1] of this lane
                    for (int c0 = 0; c0 < colEnd; c0 += 32) {
                        if (warp < 3) {
                            uint32_t e[32];
                            tmem_ld32(acc + (uint32_t)(oyl * p.CW + c0), e);
                            if (warp == 2) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(st + j) = make_float4(__uint_as_float(e[j]), __uint_as_float(e[j + 1]),
                                                                                     __uint_as_float(e[j + 2]), __uint_as_float(e[j + 3]));
                            } else if (warp == 1) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(st + j) = make_float4(j == 0 ? prev1 : __uint_as_float(e[j - 1]), __uint_as_float(e[j]),
                                                                                     __uint_as_float(e[j + 1]), __uint_as_float(e[j + 2]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(st + j) = make_float4(j == 0 ? prev0 : __uint_as_float(e[j - 2]),
                                                                                     j == 0 ? prev1 : __uint_as_float(e[j - 1]),
                                                                                     __uint_as_float(e[j]), __uint_as_float(e[j + 1]));
                            }
                            prev0 = __uint_as_float(e[30]);
                            prev1 = __uint_as_float(e[31]);
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");          // three planes staged
                        const int col = c0 + lane;
                        const int ox = tx * p.S + col - p.colBase;
                        if (col >= p.colBase && col < colEnd && ox < p.OW) {
                            OutT* yp = (OutT*)p.y + (((size_t)n * p.O + warp) * p.OH + oy) * (size_t)p.yPitch + ox;
                            const float* sp = stage + warp * STAGE_PITCH + lane;
#pragma unroll
                            for (int r = 0; r < 8; r++) {                       // channels warp, warp + 4, ..., warp + 28
                                st_as<OutT>(yp, sp[0] + sp[32 * STAGE_PITCH] + sp[64 * STAGE_PITCH]);
                                yp += 4 * chStep;
                                sp += 4 * STAGE_PITCH;
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");          // block stored: the staging buffer is free again
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&barAccEmpty[as]));
                continue;
            }
            const int nrows = min(128, p.O - o0);                 // valid channels of this tile
            // channel held by this lane: M = 128 -> lane 32w + l; M = 64 -> lanes 16 as .. 16 as + 15 of each warp hold rows 16w ..
            const int laneRow = p.m64 ? 16 * warp + (lane & 15) : 32 * warp + lane;
            const bool laneValid = !p.m64 || (uint32_t)(lane >> 4) == as;
            const bool mine = (p.m64 ? 16 * warp : 32 * warp) < nrows;
            float* st = stage + laneRow * STAGE_PITCH;
            const int colEnd = p.colBase + p.S;
            const size_t chStep = (size_t)p.OH * p.yPitch;
            for (int oyl = 0; oyl < p.R; oyl++) {
                const int oy = oy0 + oyl;
                if (oy >= p.OH) break;
                float carry = 0.f;
                for (int c0 = 0; c0 < colEnd; c0 += 32) {
                    if (mine) {
                        uint32_t e[32], o[32];
                        tmem_ld32(acc + (uint32_t)((2 * oyl) * p.CW + c0), e);
                        tmem_ld32(acc + (uint32_t)((2 * oyl + 1) * p.CW + c0), o);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 v;
                            v.x = __uint_as_float(e[j]) + (j == 0 ? carry : __uint_as_float(o[j - 1]));
                            v.y = __uint_as_float(e[j + 1]) + __uint_as_float(o[j]);
                            v.z = __uint_as_float(e[j + 2]) + __uint_as_float(o[j + 1]);
                            v.w = __uint_as_float(e[j + 3]) + __uint_as_float(o[j + 2]);
                            if (laneValid) *reinterpret_cast<float4*>(st + j) = v;
                        }
                        carry = __uint_as_float(o[31]);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");          // block staged by its owners
                    const int col = c0 + lane;
                    const int ox = tx * p.S + col - p.colBase;
                    if (col >= p.colBase && col < colEnd && ox < p.OW) {
                        typedef typename std::conditional<HALF, __half, float>::type OutT;
                        OutT* yp = (OutT*)p.y + (((size_t)n * p.O + o0 + warp) * p.OH + oy) * (size_t)p.yPitch + ox;
                        const float* sp = stage + warp * STAGE_PITCH + lane;
                        for (int r = warp; r < nrows; r += 4) {
                            st_as<OutT>(yp, *sp);
                            yp += 4 * chStep;
                            sp += 4 * STAGE_PITCH;
                        }
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");          // block stored: the staging buffer is free again
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&barAccEmpty[as]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmemCols) : "memory");
    }
}

constexpr int kSmemLimit3 = 224 * 1024;       // + 2 KB of static barriers stays under the 227 KB per-CTA limit

// Pick (NPX, R, accumulator stages, W slots) for a layer: minimise an estimate of the per-layer time
// max(MMA clocks, operand bytes / 40 B/clk) over the tile shapes that fit TMEM (512 columns) and shared memory.
bool plan_tc3(Tc3Params& p, int smemLimit, bool half)
{
    const int lose = half ? 8 : 4;          // columns of a tile that only serve as halo (box starts are 16-byte aligned)
    // Per-tile time model fitted on B200 (tools/run_tc3_sweep.sh): a fixed 1700 clk of pipeline hand-over, the K loop
    // (tensor core or operand stream, whichever is slower) and -- the accumulator is single-buffered -- the serial
    // epilogue at ~530 clk per 32-column block.  Double-buffered accumulators only fit tiles of <= 96 pixels and
    // measured slower than one large tile for every StyleGAN3-T layer, so they are not planned.
    double best = 1e300;
    bool found = false;
    const int tapSlots = p.kxStack ? 3 : 9;                      // W ring slots one chunk consumes
    for (int npx = 64; npx <= 224; npx += half ? 64 : 32) {
        for (int r = 1; r <= 4; r++) {
            // kxStack: one accumulator of npx columns per row, two stages side by side; otherwise an even / odd pair per row
            const int cw = p.kxStack ? npx : (npx + 4 + kAccPitchAlign - 1) / kAccPitchAlign * kAccPitchAlign;
            const int stageCols = p.kxStack ? r * cw : 2 * r * cw;
            // the last 32-column epilogue load of the last accumulator must stay inside the allocation
            const int need = p.kxStack ? 2 * stageCols : (2 * r - 1) * cw + ((cw + 31) & ~31);
            if (need > 512) continue;
            const int xGroup = (r + 2) * npx * 128;
            int wSlots = (smemLimit - 1024 - 4 * 32 * STAGE_PITCH * 4 - 2 * xGroup) / p.wSlotBytes;
            if (wSlots > kMaxWSlots) wSlots = kMaxWSlots;
            if (wSlots < 3) continue;
            const bool resident = wSlots >= tapSlots * p.kChunks;
            if (resident) wSlots = tapSlots * p.kChunks;
            const int s = npx - lose;
            const long long tiles = (long long)((p.OW + s - 1) / s) * ((p.OH + r - 1) / r);
            const double perMma = npx / 2.0 > 32.0 + npx / 4.0 ? npx / 2.0 : 32.0 + npx / 4.0;       // tools/tc_mma_bench.cu
            const double mma = 4.0 * tapSlots * r * perMma;
            const double load = ((resident ? 0.0 : (double)tapSlots * p.wLoadBytes) + (double)xGroup) / 80.0;     // operand stream at ~80 B/clk/SM (fitted; 40 over-penalised wide tiles)
            const double epi = 530.0 * r * ((p.colBase + s + 31) / 32);
            const double kloop = p.kChunks * (mma > load ? mma : load);
            // two accumulator stages (M = 64 in the lanes, kxStack in the columns): the epilogue of a tile overlaps the K loop of the next one
            const bool overlap = p.m64 || p.kxStack;
            const double cost = (double)tiles * (1700.0 + (overlap ? (kloop > epi ? kloop : epi) : kloop + epi));
            if (cost < best) {
                best = cost; found = true;
                p.NPX = npx; p.R = r; p.CW = cw; p.S = s; p.accStages = overlap ? 2 : 1; p.accStageCols = stageCols;
                p.wSlots = wSlots; p.xGroupBytes = xGroup; p.resident = resident ? 1 : 0;
                int cols = 32;
                while (cols < need) cols <<= 1;
                p.tmemCols = cols;
            }
        }
    }
    return found;
}

}  // namespace

// 0 when the tensor-core path takes this shape (the caller then prepares the weights tap-major), SG3_E_NOKERNEL otherwise.
int sg3_modconv_tc3_supported(int I, int O, int H, int W, int k, int pad)
{
    if (k != 3 || (pad != 0 && pad != 2)) return SG3_E_NOKERNEL;
    if (W % 4 != 0 || H + 2 * pad - 2 < 1 || W + 2 * pad - 2 < 1) return SG3_E_NOKERNEL;      // TMA: 16-byte row pitch
    (void)I; (void)O;
    return 0;
}

// x [N][I][H][xPitch >= W]; wtap [N][9][O][ldw] (tap = ky * 3 + kx, i contiguous, ldw >= I); y [N][O][OH][yPitch >= OW].
// HALF: all three are fp16.  Pitches / ldw in elements, 16-byte multiples (4 floats, 8 halves).
// The kernel only sees x through its tensor map, so a padded input row pitch costs nothing: TMA needs the PITCH to be a
// 16-byte multiple, not W (columns >= W are outside the map's extent and read as zero like any other out-of-image column).
namespace {

template <bool HALF>
int launch_tc3(const void* x, const void* wtap, void* y, int N, int I, int O, int H, int W, int pad, int ldw,
               int xPitch, int yPitch, cudaStream_t stream)
{
    constexpr int esz = HALF ? 2 : 4, gran = 16 / esz, bkc = HALF ? 64 : BK3;
    const int xp = xPitch > 0 ? xPitch : W;
    if (xp < W || xp % gran != 0) return SG3_E_NOKERNEL;
    if (pad != 0 && pad != 2) return SG3_E_NOKERNEL;
    if (H + 2 * pad - 2 < 1 || W + 2 * pad - 2 < 1) return SG3_E_NOKERNEL;
    if (ldw % gran != 0 || ldw < I) return SG3_E_NOKERNEL;
    if (((uintptr_t)x & 15) || ((uintptr_t)wtap & 15)) return SG3_E_NOKERNEL;
    Tc3Params p;
    p.y = y; p.N = N; p.I = I; p.O = O; p.H = H; p.W = W; p.pad = pad;
    p.OH = H + 2 * pad - 2; p.OW = W + 2 * pad - 2;
    p.yPitch = yPitch > 0 ? yPitch : p.OW;
    if (p.yPitch < p.OW) return SG3_E_INVALID;
    p.kChunks = (I + bkc - 1) / bkc;
    p.xoff = pad == 2 ? (HALF ? 8 : 4) : 0;
    p.colBase = p.xoff + (pad == 2 ? 0 : 2);
#ifdef SG3_NO_KXSTACK
    p.kxStack = 0;                                      // tuning / A-B builds
#else
    p.kxStack = O == 32 ? 1 : 0;                        // taps 3 ky .. 3 ky + 2 of the tap-major weights are 96 consecutive rows
#endif
    p.m64 = (!p.kxStack && O <= 64) ? 1 : 0;
    p.wRows = p.kxStack ? 32 : (O >= 128 ? 128 : (O + 7) & ~7);
    // [rows][one 128-byte K row].  kxStack: 96 rows per slot; the M = 128 MMA reads 32 rows past them (the next slot, or the staging
    // buffer behind the ring: inside the allocation), which only feed accumulator rows 96 .. 127 that nobody reads
    p.wSlotBytes = p.kxStack ? 96 * 128 : p.wRows * 128;
    p.wLoadBytes = p.wSlotBytes;
    // co-scheduling with the stencil kernel (sg3_modconv_set_smem_budget): plan inside a smaller shared-memory budget
    const int budget = sg3_conv_smem_budget();
    // (a budget too small for any tile plan is ignored: the kernel then simply does not share its SM)
    if (!(budget > 0 && budget < kSmemLimit3 && plan_tc3(p, budget, HALF)) && !plan_tc3(p, kSmemLimit3, HALF)) return SG3_E_NOKERNEL;
    p.tilesX = (p.OW + p.S - 1) / p.S;
    p.tilesY = (p.OH + p.R - 1) / p.R;
    p.tilesO = (O + 127) / 128;
    p.totalTiles = (long long)N * p.tilesY * p.tilesX * p.tilesO;
    const long long ctas = p.totalTiles < sg3_sm_count() ? p.totalTiles : sg3_sm_count();

    alignas(64) CUtensorMap mapX, mapW;
    const CUtensorMapDataType dt = HALF ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    {
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)I, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)xp * esz, (uint64_t)xp * H * esz, (uint64_t)xp * H * I * esz};
        const uint32_t box[4] = {HALF ? 64u : 32u, 1, (uint32_t)bkc, 1};
        if (!sg3_make_tensor_map(&mapX, dt, 4, x, dims, strides, box, HALF ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return SG3_E_NOKERNEL;
    }
    {
        const uint64_t dims[4] = {(uint64_t)I, (uint64_t)O, 9, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)ldw * esz, (uint64_t)ldw * O * esz, (uint64_t)ldw * O * 9 * esz};
        const uint32_t box[4] = {(uint32_t)bkc, (uint32_t)p.wRows, p.kxStack ? 3u : 1u, 1};
        if (!sg3_make_tensor_map(&mapW, dt, 4, wtap, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return SG3_E_NOKERNEL;
    }
    const int smemBytes = 2 * p.xGroupBytes + p.wSlots * p.wSlotBytes + 4 * 32 * STAGE_PITCH * 4 + 1024;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(modconv_tc3_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        // same carve-out as the stencil kernels, so that both can be resident on one SM (see modconv_tc.cu)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(modconv_tc3_kernel<HALF>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
    });
    if (attrErr != cudaSuccess) return (int)attrErr;
    modconv_tc3_kernel<HALF><<<(unsigned)ctas, kThreads3, smemBytes, stream>>>(mapX, mapW, p);
    return sg3_launch_status();
}

}  // namespace

int sg3_modconv_fwd_tc3(const float* x, const float* wtap, float* y, int N, int I, int O, int H, int W, int pad, int ldw,
                        int xPitch, int yPitch, cudaStream_t stream)
{
    return launch_tc3<false>(x, wtap, y, N, I, O, H, W, pad, ldw, xPitch, yPitch, stream);
}

int sg3_modconv_fwd_tc3_f16(const void* x, const void* wtap, void* y, int N, int I, int O, int H, int W, int pad, int ldw,
                            int xPitch, int yPitch, cudaStream_t stream)
{
    return launch_tc3<true>(x, wtap, y, N, I, O, H, W, pad, ldw, xPitch, yPitch, stream);
}
