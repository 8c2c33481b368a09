// modconv_tc.cu -- TF32 tcgen05/TMEM GEMM for the contraction of modulated_conv2d (1x1 kernels,
// StyleGAN3 config R; the grouped cuDNN convolution of networks_stylegan3.py:59-62).
//
// Per sample n:  Y[o][p] = sum_i Wn[o][i] * X[i][p]     X: [I][P] (pixels contiguous), Wn: [O][ldw]
// mapped onto UMMA as  D[M = 128 pixels][N = BN out-channels] += A[M][K] * B[K][N]  with
//   A = X^T  : MN-major (pixels contiguous), four TMA boxes [32 k-rows][32 px] per stage,
//              SWIZZLE_128B_ATOM_32B (the layout tcgen05 requires for 32-bit MN-major operands)
//   B = Wn   : K-major  (i contiguous),      one TMA box  [BN rows][32 k],      SWIZZLE_128B
//   D in TMEM: lane = pixel, column = out-channel, fp32
// so the epilogue (tcgen05.ld 32x32b: one thread = one pixel, 32 channels per load) stores, for a fixed
// channel, 32 consecutive pixels per warp instruction = one full 128-byte line; no smem staging.
// K is consumed 32 input channels per pipeline stage (4 x tcgen05.mma K=8); out-of-range rows/columns of
// both operands are zero-filled by TMA, so I, O and P need no padding in the activations (the weight
// rows are padded to `ldw` floats by the prologue so that their pitch is 16-byte aligned).
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..32w+31), warp 4 TMA producer,
// warp 5 TMEM allocator + MMA issuer.  Persistent: one CTA per SM loops over tiles; the TMA ring runs ahead across
// tiles and the accumulator is double-buffered in TMEM, so loads, MMAs and the epilogue of consecutive tiles overlap.
// Every mbarrier wait is bounded (trap instead of hang).
//
// X3 (3xTF32, "fp32x3" math): fp32-accurate contraction on the tensor cores.  Both operands are split into a TF32 head and
// a TF32 tail, W = Wh + Wl (the weight prologue writes the two planes), X = Xh + Xl with Xh = what the tensor core reads of an
// fp32 word (its top 19 bits) and Xl = X - Xh computed by four extra "splitter" warps from the landed TMA tile into a second
// smem tile; each K step issues Xh*Wh + Xh*Wl + Xl*Wh (the dropped Xl*Wl term is 2^-22 relative).  The 1x1 convs of config R
// are HBM-bound with the tensor pipe ~25 % busy, so the two extra MMAs per step mostly hide under the operand stream.
#include <cuda.h>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"
#include "tensor_map.h"

namespace {

constexpr int BM = 128;            // pixels per tile (UMMA M)
constexpr int BK = 32;             // input channels per stage (one 128-byte swizzle row of the K-major operand)
constexpr int kMaxStages = 6;
constexpr int A_STAGE_BYTES = BM * BK * 4;      // 16 KB: 4 boxes of [32 rows][128 B]
constexpr int kThreads = 192;
constexpr int kThreadsX3 = 320;   // + four splitter warps

struct TcParams {
    void* y;               // float (HALF = false) or __half (HALF = true)
    int N, I, O, P;
    int BN;                // out-channels per tile (multiple of 16, <= 256)
    int tmemCols;          // power of two >= max(BN, 32)
    int accCols;           // TMEM columns per accumulator stage (BN rounded up to 32); two stages are allocated
    int tilesM, tilesN;
    int kTiles;
    int stages;            // smem pipeline depth (4 for 48 KB stages, 6 for <= 32 KB stages)
    long long totalTiles;
    SG3_TRACE_FIELD
};

// Persistent kernel: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA producer runs
// ahead across tile boundaries (the smem ring never drains), the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
// HALF: fp16 activations / weights / output with kind::f16 MMAs.  Same stage bytes (a stage is 64 input channels of
// 2 bytes instead of 32 of 4): A = X^T is MN-major in the standard SWIZZLE_128B form (two TMA boxes [64 k][64 px] per
// stage, 8-k atoms 1024 B apart, 64-pixel blocks one box apart), B = Wn K-major as before; K = 16 per instruction.
__device__ __forceinline__ float tf32_rna(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// explicit global stores for the epilogue's walking pointers (an opaque pointer update would otherwise demote them to generic ST)
#ifndef SG3_CONV_ST
#define SG3_CONV_ST "st.global"          // tuning: e.g. "st.global.cs" (streaming stores)
#endif
__device__ __forceinline__ void st_plane(float* q, float v) { asm volatile(SG3_CONV_ST ".f32 [%0], %1;" ::"l"(q), "f"(v) : "memory"); }
__device__ __forceinline__ void st_plane(__half* q, float v)
{
    asm volatile("st.global.b16 [%0], %1;" ::"l"(q), "h"(__half_as_ushort(__float2half_rn(v))) : "memory");
}

template <bool HALF, bool X3>
__global__ void __launch_bounds__(X3 ? kThreadsX3 : kThreads, 1)
modconv_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const TcParams p)
{
    static_assert(!(HALF && X3), "the operand split is an fp32 mode");
    constexpr int KST = HALF ? 64 : 32;            // input channels per stage
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barFull[kMaxStages], barEmpty[kMaxStages], barSplit[kMaxStages], barAccFull[2], barAccEmpty[2];
    __shared__ uint32_t tmemBase;

    // warp index through a shuffle: the role dispatch is then provably warp-uniform and the producer / MMA loops run on the
    // uniform datapath (see tc_common.cuh, "warp-uniform issue")
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) SG3_TRACE_EVENT(p, 3, blockIdx.x);
    // stage layout: A (X^T tile) [| A tail (X3)] | B (W tile) [| B tail (X3)]
    constexpr int A_BYTES = X3 ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
    const int bBytes = p.BN * BK * 4;
    const int stageBytes = A_BYTES + (X3 ? 2 : 1) * bBytes;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;        // SWIZZLE_128B operands need 1024-byte aligned tiles

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; s++) {
            mbar_init(smem_u32(&barFull[s]), 1); mbar_init(smem_u32(&barEmpty[s]), 1); mbar_init(smem_u32(&barSplit[s]), 4);
        }
        for (int a = 0; a < 2; a++) { mbar_init(smem_u32(&barAccFull[a]), 1); mbar_init(smem_u32(&barAccEmpty[a]), 4); }
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

    // tile -> (sample n, pixel block tm, channel block tn); channel block fastest so that the CTAs working on one
    // pixel block at the same time share its activation tiles in L2
    auto decode = [&](long long t, int& n, int& p0, int& o0) {
        const int tn = (int)(t % p.tilesN);
        const long long rest = t / p.tilesN;
        const int tm = (int)(rest % p.tilesM);
        n = (int)(rest / p.tilesM);
        p0 = tm * BM;
        o0 = tn * p.BN;
    };

    if (warp == 4) {
        // ---------------- TMA producer (whole warp, elected lane issues) ----------------
        uint32_t it = 0;                                          // k-iterations issued so far, across tiles
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x) {
            int n, p0, o0;
            decode(t, n, p0, o0);
            for (int kt = 0; kt < p.kTiles; kt++, it++) {
                const uint32_t s = it % p.stages, round = it / p.stages;
                if (round > 0) mbar_wait(smem_u32(&barEmpty[s]), (round - 1) & 1);
                const uint32_t full = smem_u32(&barFull[s]);
                mbar_expect_tx_elect(full, (uint32_t)(A_STAGE_BYTES + (X3 ? 2 : 1) * bBytes));
                const uint32_t aDst = tiles + s * stageBytes;
                if (HALF) {
#pragma unroll
                    for (int j = 0; j < 2; j++) tma_load_3d_elect(aDst + j * (KST * 128), &mapX, full, p0 + 64 * j, kt * KST, n);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) tma_load_3d_elect(aDst + j * (BK * 128), &mapX, full, p0 + 32 * j, kt * BK, n);
                }
                if (X3) {       // weight planes [N][2][O][ldw]: head and tail
                    tma_load_3d_elect(aDst + A_BYTES, &mapW, full, kt * KST, o0, 2 * n);
                    tma_load_3d_elect(aDst + A_BYTES + bBytes, &mapW, full, kt * KST, o0, 2 * n + 1);
                } else {
                    tma_load_3d_elect(aDst + A_BYTES, &mapW, full, kt * KST, o0, n);
                }
            }
        }
    } else if (X3 && warp >= 6) {
        // ---------------- splitter (X3): Xl = X - (top 19 bits of X), elementwise on the landed tile (swizzle-agnostic) ------
        unsigned char* gen = smem + (tiles - smem_u32(smem));
        const int tid = (int)threadIdx.x - 192;                   // 0 .. 127
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x) {
            for (int kt = 0; kt < p.kTiles; kt++, it++) {
                const uint32_t s = it % p.stages;
                mbar_wait(smem_u32(&barFull[s]), (it / p.stages) & 1);
                const float4* a = (const float4*)(gen + (size_t)s * stageBytes);
                float4* lo = (float4*)(gen + (size_t)s * stageBytes + A_STAGE_BYTES);
#pragma unroll
                for (int e = 0; e < A_STAGE_BYTES / 16 / 128; e++) {
                    const float4 v = a[tid + 128 * e];
                    // the residual is exact in fp32; it is then ROUNDED to TF32 here (the tensor core would truncate it, a bias
                    // that grows linearly with the number of input channels)
                    float4 r;
                    r.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
                    r.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
                    r.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
                    r.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                    lo[tid + 128 * e] = r;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&barSplit[s]));
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer (whole warp, elected lane issues) ----------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, A MN-major, B K-major, N, M=128
        // (HALF: A = B = F16 -> format fields 0)
        const uint32_t idesc = (1u << 4) | (HALF ? 0u : (2u << 7) | (2u << 10)) | (1u << 15) | (0u << 16) |
                               ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // A (fp32, MN-major): 8 k-rows of 128 B per step = two 4-row swizzle atoms 512 B apart (SBO), 32-pixel blocks
        // BK*128 B apart (LBO), K step 1024 B;  B (K-major): K step 32 B inside the 128-byte row, 8-row groups 1024 B apart
        // A (fp16, MN-major): 16 k-rows per step = two 8-row atoms 1024 B apart (SBO), 64-pixel blocks KST*128 B apart (LBO)
        const uint64_t dA = HALF ? umma_desc(tiles, KST * 128, 1024)
                                 : umma_desc(tiles, BK * 128, 512, kLayoutSw128Base32);
        const uint64_t dB = umma_desc(tiles + A_BYTES, 16, 1024);
        const uint32_t aLo0 = (uint32_t)dA, aHi = (uint32_t)(dA >> 32), bLo0 = (uint32_t)dB, bHi = (uint32_t)(dB >> 32);
        const uint32_t stageStep = (uint32_t)stageBytes >> 4;
        uint32_t it = 0, tc = 0;                                  // k-iterations / tiles consumed so far
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x, tc++) {
            const uint32_t as = tc & 1, use = tc >> 1;            // accumulator stage and how often it was used before
            if (use > 0) mbar_wait(smem_u32(&barAccEmpty[as]), (use - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + as * p.accCols;
            for (int kt = 0; kt < p.kTiles; kt++, it++) {
                const uint32_t s = it % p.stages;
                mbar_wait(smem_u32(&barFull[s]), (it / p.stages) & 1);
                if (X3) mbar_wait(smem_u32(&barSplit[s]), (it / p.stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (HALF) umma_f16_x4<128, 2>(acc, aLo0 + s * stageStep, aHi, bLo0 + s * stageStep, bHi, idesc, kt > 0 ? 1u : 0u);
                else umma_tf32_x4<64, 2>(acc, aLo0 + s * stageStep, aHi, bLo0 + s * stageStep, bHi, idesc, kt > 0 ? 1u : 0u);
                if (X3) {
                    // + Xh * Wl + Xl * Wh  (descriptor start fields in 16-byte units)
                    umma_tf32_x4<64, 2>(acc, aLo0 + s * stageStep, aHi, bLo0 + s * stageStep + ((uint32_t)bBytes >> 4), bHi, idesc, 1u);
                    umma_tf32_x4<64, 2>(acc, aLo0 + s * stageStep + (A_STAGE_BYTES >> 4), aHi, bLo0 + s * stageStep, bHi, idesc, 1u);
                }
                umma_commit_elect(smem_u32(&barEmpty[s]));        // frees the smem stage when these MMAs retire
            }
            umma_commit_elect(smem_u32(&barAccFull[as]));         // accumulator of this tile complete
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> global (warps 0-3, TMEM lanes 32*warp .. +31) ----------------
        uint32_t tc = 0;
        for (long long t = blockIdx.x; t < p.totalTiles; t += gridDim.x, tc++) {
            int n, p0, o0;
            decode(t, n, p0, o0);
            const uint32_t as = tc & 1;
            mbar_wait(smem_u32(&barAccFull[as]), (tc >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int pix = p0 + 32 * warp + lane;
            typedef typename std::conditional<HALF, __half, float>::type OutT;
            // The store loop is the critical path of the HBM-bound layers: ONE warp owns 32 pixels of a tile and has to issue every
            // store of their BN channels itself.  With the address (64-bit multiply) and two bound checks evaluated per element this
            // was ~19 instructions per store -- 2000 dependent instructions per warp and tile, and the ncu source view showed the MMA
            // and TMA warps asleep on their barriers while the epilogue warps crawled (L13: 0.64 of HBM).  Now: one pointer per
            // 32-channel block walking down the channel planes, the bound check hoisted to a warp-uniform block count.
            const size_t planeStep = (size_t)p.P;
            OutT* yb = (OutT*)p.y + ((size_t)n * p.O + o0) * planeStep + pix;          // channel o0, this thread's pixel
            const bool pixOk = pix < p.P;
            for (int c = 0; c < p.BN; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + as * p.accCols + (uint32_t)c, r);
                const int nvalid = min(p.BN - c, p.O - o0 - c);                          // channels of this block that exist (uniform)
                if (pixOk) {
                    // two pointers (even / odd channels) walking down the planes; the empty asm keeps ptxas from turning the walk
                    // back into 32 independent 64-bit multiplies (which it hoists, at 64 registers and ~400 IMADs per tile)
                    OutT* q0 = yb + (size_t)c * planeStep;
                    OutT* q1 = q0 + planeStep;
                    const size_t step2 = 2 * planeStep;
                    if (nvalid >= 32) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            st_plane(q0, __uint_as_float(r[j]));
                            st_plane(q1, __uint_as_float(r[j + 1]));
                            q0 += step2; q1 += step2;
                            asm volatile("" : "+l"(q0), "+l"(q1));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            if (j < nvalid) st_plane(q0, __uint_as_float(r[j]));
                            if (j + 1 < nvalid) st_plane(q1, __uint_as_float(r[j + 1]));
                            q0 += step2; q1 += step2;
                            asm volatile("" : "+l"(q0), "+l"(q1));
                        }
                    }
                }
            }
            // this warp's TMEM reads are done: hand the accumulator stage back to the MMA warp
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
        if (lane == 0) SG3_TRACE_EVENT(p, 4, blockIdx.x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad of the 1x1 modulated conv:  dW[n][o][i] = sum_p dY[n][o][p] * X[n][i][p]   (per sample; networks_stylegan3.py:59-62
// differentiated wrt the modulated weights).  D[M = 128 out-channels][N = BN in-channels] += A[M][K] * B[K][N] with both
// operands K-major (pixels contiguous): A = dY tile, B = X tile, K = pixels, 32 per stage.  The pixel range is split across
// CTAs (split-K); every CTA adds its partial tile into dW with fp32 atomics (dW is zeroed by the caller).
struct WgParams {
    float* dw;             // [N][O][ldw]
    int N, I, O, P, ldw;
    int BN, tmemCols;
    int tilesO, tilesI, splits, kPerSplit;   // k-tiles (32 pixels) per split
    int kTilesTotal;
    int stages;            // operand ring depth
    int vec;               // dw rows are 16-byte aligned (ldw % 4 == 0, aligned base): 128-bit reductions
};

__global__ void __launch_bounds__(kThreads, 1)
modconv_wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX, const WgParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barFull[kMaxStages], barEmpty[kMaxStages], barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int stageBytes = A_STAGE_BYTES + p.BN * BK * 4;
    const int stages = p.stages;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;

    long long t = blockIdx.x;
    const int sp = (int)(t % p.splits); t /= p.splits;
    const int ti = (int)(t % p.tilesI); t /= p.tilesI;
    const int to = (int)(t % p.tilesO);
    const int n = (int)(t / p.tilesO);
    const int o0 = to * BM, i0 = ti * p.BN;
    // contiguous k-tile range per split.  (Measured: interleaving the splits -- tiles sp, sp + splits, ... so that co-running CTAs read
    // adjacent pieces of every channel row -- changes nothing: 4.2 TB/s on the top layers either way, tools/prof_wgrad1.py.)
    const int kt0 = sp * p.kPerSplit, ktStep = 1;
    const int nk = min(kt0 + p.kPerSplit, p.kTilesTotal) - kt0;          // >= 1 by construction

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; s++) { mbar_init(smem_u32(&barFull[s]), 1); mbar_init(smem_u32(&barEmpty[s]), 1); }
        mbar_init(smem_u32(&barAccum), 1);
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

    if (warp == 4) {
        for (int it = 0; it < nk; it++) {
            const int s = it % stages, round = it / stages;
            if (round > 0) mbar_wait(smem_u32(&barEmpty[s]), (round - 1) & 1);
            const uint32_t full = smem_u32(&barFull[s]);
            mbar_expect_tx_elect(full, (uint32_t)stageBytes);
            const uint32_t aDst = tiles + s * stageBytes;
            tma_load_3d_elect(aDst, &mapDY, full, (kt0 + it * ktStep) * BK, o0, n);                    // [128 o][32 px]
            tma_load_3d_elect(aDst + A_STAGE_BYTES, &mapX, full, (kt0 + it * ktStep) * BK, i0, n);     // [BN i][32 px]
        }
    } else if (warp == 5) {
        // D=F32, A=B=TF32, both K-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint64_t dA = umma_desc(tiles, 16, 1024), dB = umma_desc(tiles + A_STAGE_BYTES, 16, 1024);
        const uint32_t aLo0 = (uint32_t)dA, aHi = (uint32_t)(dA >> 32), bLo0 = (uint32_t)dB, bHi = (uint32_t)(dB >> 32);
        const uint32_t stageStep = (uint32_t)stageBytes >> 4;
        for (int it = 0; it < nk; it++) {
            const int s = it % stages;
            mbar_wait(smem_u32(&barFull[s]), (it / stages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_tf32_x4<2, 2>(tmem, aLo0 + s * stageStep, aHi, bLo0 + s * stageStep, bHi, idesc, it > 0 ? 1u : 0u);
            umma_commit_elect(smem_u32(&barEmpty[s]));
        }
        umma_commit_elect(smem_u32(&barAccum));
    } else {
        mbar_wait(smem_u32(&barAccum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int o = o0 + 32 * warp + lane;
        float* row = p.dw + ((size_t)n * p.O + (size_t)(o < p.O ? o : 0)) * p.ldw;
        for (int c = 0; c < p.BN; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c, r);
            if (o < p.O) {
                // 128-bit reductions (REDG.E.ADD.F32x4): rows are ldw = 4k floats apart and i0 + c is a multiple of 16; the columns in
                // [I, ldw) receive the zero products of TMA's zero-filled channel rows.  (Scalar reductions, one row per lane, were
                // most of this kernel on the low-resolution layers: measured on the 3x3 sibling, profiles/r02_wgrad3.md.)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int i = i0 + c + j;
                    if (p.vec) {
                        if (c + j < p.BN && i < p.ldw)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + i), "f"(__uint_as_float(r[j])),
                                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (c + j + q < p.BN && i + q < p.I) atomicAdd(row + i + q, __uint_as_float(r[j + q]));
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmemCols) : "memory");
    }
}

// ---- host: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    });
    return fn;
}

}  // namespace

// Generic tiled tensor-map encoder shared with the filtered_lrelu TMA path.
bool sg3_make_tensor_map(CUtensorMap* m, CUtensorMapDataType type, int rank, const void* base, const uint64_t* dims,
                         const uint64_t* stridesBytes, const uint32_t* box, CUtensorMapSwizzle swizzle)
{
    EncodeTiledFn enc = get_encode();
    if (!enc || rank < 1 || rank > 5) return false;
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) gs[i] = stridesBytes[i];
#ifndef SG3_TMA_PROMO
#define SG3_TMA_PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B      // tuning: L2 promotion size of every tensor map
#endif
    return enc(m, type, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
               SG3_TMA_PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

bool make_map3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1Bytes, uint64_t s2Bytes,
               uint32_t b0, uint32_t b1, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1Bytes, s2Bytes};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

namespace {

// x [N][I][P], wmod [N][O][ldw] (zero beyond I), y [N][O][P]; k must be 1.  HALF: all three are fp16.
template <bool HALF, bool X3>
int launch_fwd_tc(const void* x, const void* wmod, void* y, int N, int I, int O, int H, int W, int k, int pad, int ldw, cudaStream_t stream)
{
    if (k != 1 || pad != 0) return SG3_E_NOKERNEL;
    const int esz = HALF ? 2 : 4, kst = HALF ? 64 : 32;
    const long long P = (long long)H * W;
    if ((P * esz) % 16 != 0 || (ldw * esz) % 16 != 0 || ldw < I) return SG3_E_NOKERNEL;        // TMA needs 16-byte global strides
    if (((uintptr_t)x & 15) || ((uintptr_t)wmod & 15)) return SG3_E_NOKERNEL;
    if (P > INT32_MAX) return SG3_E_TOOLARGE;

    TcParams p;
    p.y = y; p.N = N; p.I = I; p.O = O; p.P = (int)P;
    // out-channel tile: split O evenly into the fewest tiles of <= 256, rounded up to the UMMA N granule (16)
    constexpr int kMaxBN = X3 ? 128 : 256;                  // X3 stages hold two A tiles and two B tiles: 64 KB at BN = 128
    const int nt = (O + kMaxBN - 1) / kMaxBN;
    int bn = ((O + nt - 1) / nt + 15) & ~15;
    if (bn < 16) bn = 16;
    p.BN = bn;
    p.tilesN = (O + bn - 1) / bn;
    p.tilesM = (int)((P + BM - 1) / BM);
    p.kTiles = (I + kst - 1) / kst;
    p.accCols = (bn + 31) & ~31;
    int cols = 32;
    while (cols < 2 * p.accCols) cols <<= 1;                 // two accumulator stages, power-of-two allocation (<= 512)
    p.tmemCols = cols;
    p.totalTiles = (long long)N * p.tilesM * p.tilesN;
    const long long ctas = p.totalTiles < sg3_sm_count() ? p.totalTiles : sg3_sm_count();   // persistent: one CTA per SM

    alignas(64) CUtensorMap mapX, mapW;
    const CUtensorMapDataType dt = HALF ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    {
        const uint64_t dims[3] = {(uint64_t)P, (uint64_t)I, (uint64_t)N};
        const uint64_t strides[2] = {(uint64_t)P * esz, (uint64_t)P * I * esz};
        const uint32_t box[3] = {HALF ? 64u : 32u, (uint32_t)kst, 1};
        if (!sg3_make_tensor_map(&mapX, dt, 3, x, dims, strides, box, HALF ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return SG3_E_NOKERNEL;
    }
    {
        // X3: the weight tensor is [N][2][O][ldw] (TF32 head and tail planes): third dimension 2N
        const uint64_t dims[3] = {(uint64_t)ldw, (uint64_t)O, (uint64_t)(X3 ? 2 * N : N)};
        const uint64_t strides[2] = {(uint64_t)ldw * esz, (uint64_t)ldw * O * esz};
        const uint32_t box[3] = {(uint32_t)kst, (uint32_t)bn, 1};
        if (!sg3_make_tensor_map(&mapW, dt, 3, wmod, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SG3_E_NOKERNEL;
    }
    // operand stages: 4 x 48 KB or 6 x <= 32 KB; X3: 3 x <= 64 KB.  Measured (tools/prof_conv.py 32 R, alternating A/B runs): a deeper
    // ring for the narrow top layers (8 x 24 KB at BN = 64, 10 x 18 KB at BN = 16) is SLOWER (L12 4.17 -> 4.58 ms, L13 / L14 unchanged):
    // more row streams in flight in DRAM, as for the stencil kernels.
    p.stages = X3 ? 3 : (bn > 128 ? 4 : 6);
    SG3_TRACE_SET(p);
    const int stageBytes = (X3 ? 2 : 1) * (A_STAGE_BYTES + bn * BK * 4);
    if (const int budget = sg3_conv_smem_budget()) {
        // co-scheduling with the stencil kernel (sg3_modconv_set_smem_budget): a shallower ring so that this CTA fits NEXT TO three
        // resident stencil CTAs; the contraction is HBM-bound and 2-4 stages of 32-48 KB still cover the DRAM latency
        const int fit = (budget - 1024) / stageBytes;
        if (fit < p.stages) p.stages = fit < 2 ? 2 : fit;
    }
    const int smemBytes = p.stages * stageBytes + 1024;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(modconv_tc_kernel<HALF, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        // the same (maximal) shared-memory carve-out as the stencil kernels: an SM cannot hold CTAs of two kernels that were
        // launched with different carve-outs, and PipelinedSynthesis wants this CTA next to three stencil CTAs
        if (e == cudaSuccess) e = cudaFuncSetAttribute(modconv_tc_kernel<HALF, X3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
    });
    if (attrErr != cudaSuccess) return (int)attrErr;
    modconv_tc_kernel<HALF, X3><<<(unsigned)ctas, X3 ? kThreadsX3 : kThreads, smemBytes, stream>>>(mapX, mapW, p);
    return sg3_launch_status();
}

}  // namespace

int sg3_modconv_fwd_tc(const float* x, const float* wmod, float* y, int N, int I, int O, int H, int W, int k, int pad, int ldw,
                       cudaStream_t stream)
{
    return launch_fwd_tc<false, false>(x, wmod, y, N, I, O, H, W, k, pad, ldw, stream);
}

// 3xTF32: wmod is [N][2][O][ldw] (sg3_modconv_weights with round_tf32 = 3)
int sg3_modconv_fwd_tc_x3(const float* x, const float* wmod, float* y, int N, int I, int O, int H, int W, int k, int pad, int ldw,
                          cudaStream_t stream)
{
    return launch_fwd_tc<false, true>(x, wmod, y, N, I, O, H, W, k, pad, ldw, stream);
}

int sg3_modconv_fwd_tc_f16(const void* x, const void* wmod, void* y, int N, int I, int O, int H, int W, int k, int pad, int ldw,
                           cudaStream_t stream)
{
    return launch_fwd_tc<true, false>(x, wmod, y, N, I, O, H, W, k, pad, ldw, stream);
}

// dy [N][O][P], x [N][I][P], dw [N][O][ldw] (zeroed by the caller; ldw >= I).
int sg3_modconv_wgrad_tc(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int ldw, cudaStream_t stream)
{
    const long long P = (long long)H * W;
    if (P % 4 != 0 || P > INT32_MAX) return SG3_E_NOKERNEL;
    if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15)) return SG3_E_NOKERNEL;
    WgParams p;
    p.dw = dw; p.N = N; p.I = I; p.O = O; p.P = (int)P; p.ldw = ldw;
    p.vec = (ldw % 4 == 0 && ((uintptr_t)dw & 15) == 0) ? 1 : 0;
    const int nt = (I + 255) / 256;
    int bn = ((I + nt - 1) / nt + 15) & ~15;
    if (bn < 16) bn = 16;
    p.BN = bn;
    p.tilesI = (I + bn - 1) / bn;
    p.tilesO = (O + BM - 1) / BM;
    int cols = 32;
    while (cols < bn) cols <<= 1;
    p.tmemCols = cols;
    p.kTilesTotal = (int)((P + BK - 1) / BK);
    // split the pixel range so that a few waves of CTAs exist; at least 8 k-tiles per CTA
    const long long baseTiles = (long long)N * p.tilesO * p.tilesI;
    long long splits = ((long long)sg3_sm_count() * 2 + baseTiles - 1) / baseTiles;
    const long long maxSplits = (p.kTilesTotal + 7) / 8;
    if (splits > maxSplits) splits = maxSplits;
    if (splits < 1) splits = 1;
    p.kPerSplit = (int)((p.kTilesTotal + splits - 1) / splits);
    p.splits = (p.kTilesTotal + p.kPerSplit - 1) / p.kPerSplit;
    const long long ctas = baseTiles * p.splits;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;

    alignas(64) CUtensorMap mapDY, mapX;
    if (!make_map3(&mapDY, dy, (uint64_t)P, (uint64_t)O, (uint64_t)N, (uint64_t)P * 4, (uint64_t)P * O * 4, BK, BM)) return SG3_E_NOKERNEL;
    if (!make_map3(&mapX, x, (uint64_t)P, (uint64_t)I, (uint64_t)N, (uint64_t)P * 4, (uint64_t)P * I * 4, BK, (uint32_t)bn)) return SG3_E_NOKERNEL;
    const int stages = bn > 128 ? 4 : 6;
    p.stages = stages;
    const int smemBytes = stages * (A_STAGE_BYTES + bn * BK * 4) + 1024;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([] { return cudaFuncSetAttribute(modconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attrErr != cudaSuccess) return (int)attrErr;
    modconv_wgrad_kernel<<<(unsigned)ctas, kThreads, smemBytes, stream>>>(mapDY, mapX, p);
    return sg3_launch_status();
}
