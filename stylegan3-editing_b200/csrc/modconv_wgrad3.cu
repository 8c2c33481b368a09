// modconv_wgrad3.cu -- TF32 tcgen05/TMEM weight gradient of the 3x3 modulated conv (StyleGAN3 config T; the grouped
// convolution of networks_stylegan3.py:59-62 differentiated wrt the per-sample weights; conv2d_gradfix.py:153-174 is what the
// reference runs for it).
//
//   dW[n][ky][kx][o][i] = sum_oy sum_ox  dY[n][o][oy][ox] * X[n][i][oy + ky - pad][ox + kx - pad]        (X = 0 outside the image)
//
// One GEMM per tap with K = output pixels:  D_tap[M = 128 o][N = BN i] += A[M][K] * B_tap[K][N], both operands K-major (pixels
// contiguous), 32 pixels of one image row per K chunk.  A = dY tile [128 o][32 px] lands by TMA (SWIZZLE_128B).  B_tap needs X
// shifted by kx - pad pixels, which is neither a legal TMA box start nor a legal UMMA descriptor start (16-byte granularity
// both), so the shift is done by the four otherwise idle epilogue warps: TMA lands an unswizzled [BN i][36 px] window of one
// X row (start 4-pixel aligned, zero fill outside the image = the conv padding), the warps write the three shifted copies
// [BN i][32 px] in the SWIZZLE_128B K-major form (rounded to the nearest TF32 value on the way) and hand them to the MMA warp
// through fence.proxy.async + an mbarrier.  The ky shift is a row choice: X row r pairs with dY rows r + pad - ky, so the dY
// tiles of a column chunk are kept in a ring while the kernel walks down the rows (each tile is loaded once, used by three X rows).
// All nine accumulators live in TMEM at once (9 * BN <= 512 columns -> BN <= 48): one pass over X and dY per (o tile, i tile);
// the three kx copies are contiguous in shared memory, so one MMA of N = 3 BN per ky updates three of them.
// The row range is split across CTAs; every CTA adds its nine partial tiles into dW with fp32 atomics (dW zeroed by the caller).
//
// Warp roles (192 threads): warps 0-3 shifters, then epilogue; warp 4 TMA producer; warp 5 TMEM allocator + MMA issuer.
// Every mbarrier wait is bounded (trap instead of hang).
#include <cuda.h>

#include "common.cuh"
#include "tensor_map.h"
#include "tc_common.cuh"

namespace {

constexpr int kThreadsW3 = 192;
constexpr int kDySlots = 8;                  // ring of dY tiles: three live + five in flight
constexpr int kRawStages = 4;                // X row windows in flight (the shifted copies are double-buffered)
constexpr int kDyTileBytes = 128 * 32 * 4;   // [128 o][32 px]
constexpr int kRawW = 36;                    // pixels of an X row window: nine 16-byte granules, start 4 px left of the chunk (pad 2) or at it (pad 0)

struct Wg3Params {
    float* dw;                 // [N][9][O][ldw]
    int N, I, O, H, W, OH, OW, pad, ldw;
    int BN;                    // input channels per tile (16 / 32 / 48)
    int tilesO, tilesI, splits, rowsPerSplit, nCx;
};

__device__ __forceinline__ uint32_t cvt_tf32(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__global__ void __launch_bounds__(kThreadsW3, 1)
modconv_wgrad3_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX, const Wg3Params p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barDyFull[kDySlots], barDyEmpty[kDySlots], barRawFull[kRawStages], barRawEmpty[kRawStages], barShFull[2], barShEmpty[2], barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t shTileBytes = (uint32_t)p.BN * 128u;             // one shifted copy [BN i][32 px]
    const uint32_t rawBytes = (uint32_t)p.BN * (kRawW * 4);
    const uint32_t dyRing = base;
    const uint32_t shRing = base + kDySlots * kDyTileBytes;          // 2 stages x 3 copies (1024-aligned: BN % 8 == 0)
    const uint32_t rawRing = shRing + 6u * shTileBytes;              // kRawStages windows

    long long tb = blockIdx.x;
    const int ti = (int)(tb % p.tilesI); tb /= p.tilesI;
    const int to = (int)(tb % p.tilesO); tb /= p.tilesO;
    const int sp = (int)(tb % p.splits);
    const int n = (int)(tb / p.splits);
    const int o0 = to * 128, i0 = ti * p.BN;
    const int rA = sp * p.rowsPerSplit;
    const int rows = min(p.H, rA + p.rowsPerSplit) - rA;             // >= 1 by construction

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDySlots; s++) { mbar_init(smem_u32(&barDyFull[s]), 1); mbar_init(smem_u32(&barDyEmpty[s]), 1); }
        for (int s = 0; s < kRawStages; s++) { mbar_init(smem_u32(&barRawFull[s]), 1); mbar_init(smem_u32(&barRawEmpty[s]), 4); }
        for (int s = 0; s < 2; s++) { mbar_init(smem_u32(&barShFull[s]), 4); mbar_init(smem_u32(&barShEmpty[s]), 1); }
        mbar_init(smem_u32(&barAccum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmemBase, 0);

    if (warp == 4) {
        // ---------------- TMA producer: per column chunk the dY rows rA + pad - 2 ... rB - 1 + pad, and the X rows rA ... rB - 1 ------------
        uint32_t g = 0, st = 0;                                   // dY tiles / X windows issued so far
        auto load_dy = [&](int ox0, int row) {
            const uint32_t slot = g % kDySlots, round = g / kDySlots;
            if (round > 0) mbar_wait(smem_u32(&barDyEmpty[slot]), (round - 1) & 1);
            const uint32_t full = smem_u32(&barDyFull[slot]);
            mbar_expect_tx_elect(full, kDyTileBytes);
            tma_load_4d_elect(dyRing + slot * kDyTileBytes, &mapDY, full, ox0, row, o0, n);
            g++;
        };
        for (int cx = 0; cx < p.nCx; cx++) {
            const int ox0 = cx * 32;
            load_dy(ox0, rA + p.pad - 2);
            load_dy(ox0, rA + p.pad - 1);
            for (int t = 0; t < rows; t++, st++) {
                load_dy(ox0, rA + p.pad + t);
                const uint32_t s = st % kRawStages;
                if (st >= kRawStages) mbar_wait(smem_u32(&barRawEmpty[s]), (st / kRawStages - 1) & 1);
                const uint32_t full = smem_u32(&barRawFull[s]);
                mbar_expect_tx_elect(full, rawBytes);
                tma_load_4d_elect(rawRing + s * rawBytes, &mapX, full, ox0 - 2 * p.pad, rA + t, i0, n);
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer: D = F32, A = B = TF32, both K-major, M = 128, N = 3 BN ----------------
        // The three shifted copies of a stage are contiguous [3 BN rows][32 px]: ONE MMA per ky covers kx = 0, 1, 2 (accumulator
        // columns (ky * 3 + kx) * BN + i), so the 4 KB dY operand is read once instead of three times -- with N = BN the
        // instruction is bound by its shared-memory operand reads (44 clk at N = 48 against 25 clk of math, tools/tc_mma_bench.cu).
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((3 * p.BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dA = umma_desc(dyRing, 16, 1024), dB = umma_desc(shRing, 16, 1024);
        const uint32_t aLo0 = (uint32_t)dA, aHi = (uint32_t)(dA >> 32), bLo0 = (uint32_t)dB, bHi = (uint32_t)(dB >> 32);
        const uint32_t shStep = shTileBytes >> 4;
        uint32_t gBase = 0, st = 0;
        for (int cx = 0; cx < p.nCx; cx++) {
            int waited = 0;
            for (int t = 0; t < rows + 2; t++) {
                if (t < rows) {
                    for (; waited < t + 3; waited++) {
                        const uint32_t e = gBase + (uint32_t)waited;
                        mbar_wait(smem_u32(&barDyFull[e % kDySlots]), (e / kDySlots) & 1);
                    }
                    const uint32_t s = st & 1;
                    mbar_wait(smem_u32(&barShFull[s]), (st >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t first = (cx == 0 && t == 0) ? 0u : 1u;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++) {
                        const uint32_t slot = (gBase + (uint32_t)(t + 2 - ky)) % kDySlots;       // dY row (rA + t) + pad - ky
                        umma_tf32_x4<2, 2>(tmem + (uint32_t)(ky * 3 * p.BN), aLo0 + slot * (kDyTileBytes >> 4), aHi,
                                           bLo0 + s * 3u * shStep, bHi, idesc, first);
                    }
                    umma_commit_elect(smem_u32(&barShEmpty[s]));
                    st++;
                }
                // tile t of this column was last read by step t (ky = 2); the two drain steps release the last two tiles
                umma_commit_elect(smem_u32(&barDyEmpty[(gBase + (uint32_t)t) % kDySlots]));
            }
            gBase += (uint32_t)(rows + 2);
        }
        umma_commit_elect(smem_u32(&barAccum));
    } else {
        // ---------------- shifters: X window [BN i][36 px] -> three copies [BN i][32 px], K-major SWIZZLE_128B ----------------
        // One item = (channel row i, 16-byte granule g of the chunk): two aligned 128-bit loads give window pixels 4g .. 4g + 7,
        // the three copies of the granule are three 128-bit stores (granule g of row i sits at g ^ (i & 7) in the swizzled row).
        // A quarter warp covers the eight granules of one row: loads and stores are conflict-free.  All loads of a stage are
        // issued before the first store (shared-memory stores would otherwise order the loads behind them).
        unsigned char* gen = smem + (base - smem_u32(smem));
        const unsigned char* rawG = gen + (rawRing - base);
        unsigned char* shG = gen + (shRing - base);
        const int tid = (int)threadIdx.x;                          // 0 .. 127
        const int g = tid & 7;
        const int items = p.BN >> 4;                               // BN * 8 / 128 rows-of-granules per thread: 1 .. 3
        const bool pad2 = p.pad == 2;                              // window starts 4 px left of the chunk; pad 0: at the chunk
        uint32_t st = 0;
        for (int cx = 0; cx < p.nCx; cx++) {
            for (int t = 0; t < rows; t++, st++) {
                const uint32_t s = st & 1, rs = st % kRawStages;
                mbar_wait(smem_u32(&barRawFull[rs]), (st / kRawStages) & 1);
                if (st >= 2) mbar_wait(smem_u32(&barShEmpty[s]), ((st >> 1) - 1) & 1);
                const unsigned char* raw = rawG + (size_t)rs * rawBytes;
                unsigned char* sh = shG + (size_t)s * 3u * shTileBytes;
                float4 lo[3], hi[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    if (q < items) {
                        const int i = (tid >> 3) + 16 * q;
                        const float4* rp = reinterpret_cast<const float4*>(raw + i * (kRawW * 4)) + g;
                        lo[q] = rp[0];
                        hi[q] = rp[1];
                    }
                }
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    if (q < items) {
                        const int i = (tid >> 3) + 16 * q;
                        uint32_t a[8] = {cvt_tf32(lo[q].x), cvt_tf32(lo[q].y), cvt_tf32(lo[q].z), cvt_tf32(lo[q].w),
                                         cvt_tf32(hi[q].x), cvt_tf32(hi[q].y), cvt_tf32(hi[q].z), cvt_tf32(hi[q].w)};
                        unsigned char* dst = sh + (uint32_t)i * 128u + ((uint32_t)(g ^ (i & 7)) << 4);
                        if (pad2) {                                // chunk pixel j + kx - 2 = window pixel j + kx + 2
                            *reinterpret_cast<uint4*>(dst) = make_uint4(a[2], a[3], a[4], a[5]);
                            *reinterpret_cast<uint4*>(dst + shTileBytes) = make_uint4(a[3], a[4], a[5], a[6]);
                            *reinterpret_cast<uint4*>(dst + 2 * shTileBytes) = make_uint4(a[4], a[5], a[6], a[7]);
                        } else {
                            *reinterpret_cast<uint4*>(dst) = make_uint4(a[0], a[1], a[2], a[3]);
                            *reinterpret_cast<uint4*>(dst + shTileBytes) = make_uint4(a[1], a[2], a[3], a[4]);
                            *reinterpret_cast<uint4*>(dst + 2 * shTileBytes) = make_uint4(a[2], a[3], a[4], a[5]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&barShFull[s]));
                    mbar_arrive(smem_u32(&barRawEmpty[rs]));
                }
            }
        }
        // ---------------- epilogue: nine partial tiles -> dW (fp32 atomics) ----------------
        mbar_wait(smem_u32(&barAccum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int o = o0 + 32 * warp + lane;
        for (int tap = 0; tap < 9; tap++) {
            float* row = p.dw + (((size_t)n * 9 + tap) * p.O + (size_t)(o < p.O ? o : 0)) * p.ldw;
            for (int c = 0; c < p.BN; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(tap * p.BN + c), r);
                if (o < p.O) {
                    // 128-bit reductions: four input channels per instruction (rows are ldw = 4k floats apart and i0 + c is a
                    // multiple of 16); columns in [I, ldw) receive the zero products of TMA's zero-filled channel rows
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int i = i0 + c + j;
                        if (c + j < p.BN && i < p.ldw)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + i), "f"(__uint_as_float(r[j])),
                                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

// dy [N][O][OH][dyPitch >= OW], x [N][I][H][xPitch >= W] (pitches in floats, multiples of 4; 0 = dense), OH = H + 2 pad - 2;
// dw [N][9][O][ldw >= I] (tap = ky * 3 + kx), zeroed by the caller.
int sg3_modconv_wgrad3_tc(const float* dy, const float* x, float* dw, int N, int I, int O, int H, int W, int pad, int ldw,
                          int dyPitch, int xPitch, cudaStream_t stream)
{
    if (pad != 0 && pad != 2) return SG3_E_NOKERNEL;
    const int OH = H + 2 * pad - 2, OW = W + 2 * pad - 2;
    if (OH < 1 || OW < 1) return SG3_E_NOKERNEL;
    const int xp = xPitch > 0 ? xPitch : W, dyp = dyPitch > 0 ? dyPitch : OW;
    if (xp < W || dyp < OW || xp % 4 != 0 || dyp % 4 != 0 || ldw < I) return SG3_E_NOKERNEL;        // TMA: 16-byte row pitches
    if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15)) return SG3_E_NOKERNEL;
    if (ldw % 4 != 0 || ((uintptr_t)dw & 15)) return SG3_E_NOKERNEL;                                  // 128-bit reductions into dw
    if (N < 1 || I < 1 || O < 1) return SG3_E_INVALID;

    Wg3Params p;
    p.dw = dw; p.N = N; p.I = I; p.O = O; p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.pad = pad; p.ldw = ldw;
    const int nt = (I + 47) / 48;
    p.BN = ((I + nt - 1) / nt + 15) & ~15;                  // 16 / 32 / 48
    p.tilesI = (I + p.BN - 1) / p.BN;
    p.tilesO = (O + 127) / 128;
    p.nCx = (OW + 31) / 32;
    // Split the rows across CTAs.  Every CTA holds all of TMEM, so the grid runs in waves of one CTA per SM: among the split counts
    // that give at most 6 waves pick the cheapest under a small cost model (a split walks at least 8 rows; two extra dY tiles per
    // column chunk and one more epilogue are the price of a split), preferring fewer splits on a tie.
    const long long baseTiles = (long long)N * p.tilesO * p.tilesI;
    const int sms = sg3_sm_count();
    const int maxSplits = (H + 7) / 8;
    double bestCost = 1e300;
    int bestRows = H;
    for (int sTry = 1; sTry <= maxSplits; sTry++) {
        const int rowsPer = (H + sTry - 1) / sTry;
        const long long c = baseTiles * ((H + rowsPer - 1) / rowsPer);
        const long long waves = (c + sms - 1) / sms;
        if (waves > 6 && sTry > 1) break;
        // per CTA: (rows + 2 drain steps) x column chunks x ~1000 clk per step, + the epilogue (9 BN x 128 values as 128-bit
        // reductions, ~0.4 clk per value measured on the 36 x 36 layers) + ~4000 clk of set-up
        const double cost = (double)waves * ((rowsPer + 2.0) * p.nCx * 1000.0 + 0.4 * 9 * p.BN * 128 + 4000.0);
        if (cost < bestCost * 0.999) { bestCost = cost; bestRows = rowsPer; }
    }
    p.rowsPerSplit = bestRows;
    p.splits = (H + p.rowsPerSplit - 1) / p.rowsPerSplit;
    const long long ctas = baseTiles * p.splits;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;

    alignas(64) CUtensorMap mapDY, mapX;
    {
        const uint64_t dims[4] = {(uint64_t)OW, (uint64_t)OH, (uint64_t)O, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)dyp * 4, (uint64_t)dyp * OH * 4, (uint64_t)dyp * OH * O * 4};
        const uint32_t box[4] = {32, 1, 128, 1};
        if (!sg3_make_tensor_map(&mapDY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dy, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return SG3_E_NOKERNEL;
    }
    {
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)I, (uint64_t)N};
        const uint64_t strides[3] = {(uint64_t)xp * 4, (uint64_t)xp * H * 4, (uint64_t)xp * H * I * 4};
        const uint32_t box[4] = {kRawW, 1, (uint32_t)p.BN, 1};
        if (!sg3_make_tensor_map(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return SG3_E_NOKERNEL;
    }
    const int smemBytes = kDySlots * kDyTileBytes + 6 * p.BN * 128 + kRawStages * p.BN * kRawW * 4 + 1024;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([] { return cudaFuncSetAttribute(modconv_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); });
    if (attrErr != cudaSuccess) return (int)attrErr;
    modconv_wgrad3_kernel<<<(unsigned)ctas, kThreadsW3, smemBytes, stream>>>(mapDY, mapX, p);
    return sg3_launch_status();
}
