// modconv_wgrad3.cu -- TF32 tcgen05/TMEM weight gradient of the 3x3 modulated conv (StyleGAN3 config T; the grouped
// convolution of networks_stylegan3.py:59-62 differentiated wrt the per-sample weights; conv2d_gradfix.py:103-129 is what the
// reference runs for it).
//
//   dW[n][ky][kx][o][i] = sum_oy sum_ox  dY[n][o][oy][ox] * X[n][i][oy + ky - pad][ox + kx - pad]        (X = 0 outside the image)
//
// One GEMM per tap with K = output pixels:  D_tap[M = 128 o][N = BN i] += A[M][K] * B_tap[K][N], both operands K-major (pixels
// contiguous), 32 pixels of one image row per K chunk.  A = dY tile [128 o][32 px] lands by TMA (SWIZZLE_128B).  B_tap needs X
// shifted by kx - pad pixels, which is neither a legal TMA box start nor a legal UMMA descriptor start (16-byte granularity
// both), so the shift is done by the four otherwise idle epilogue warps: TMA lands an unswizzled [BN i][40 px] window of one
// X row (start 4-pixel aligned, zero fill outside the image = the conv padding), the warps write the three shifted copies
// [BN i][32 px] in the SWIZZLE_128B K-major form (rounded to the nearest TF32 value on the way) and hand them to the MMA warp
// through fence.proxy.async + an mbarrier.  The ky shift is a row choice: X row r pairs with dY rows r + pad - ky, so the dY
// tiles of a column chunk are kept in a ring while the kernel walks down the rows (each tile is loaded once, used by three X rows).
// All nine accumulators live in TMEM at once (9 * BN <= 512 columns -> BN <= 48): one pass over X and dY per (o tile, i tile).
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
constexpr int kDySlots = 6;                  // ring of dY tiles: three live + three in flight
constexpr int kDyTileBytes = 128 * 32 * 4;   // [128 o][32 px]
constexpr int kRawW = 40;                    // pixels of an X row window: 32 + the shifts, start 4 px left of the chunk

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
    __shared__ __align__(8) uint64_t barDyFull[kDySlots], barDyEmpty[kDySlots], barRawFull[2], barRawEmpty[2], barShFull[2], barShEmpty[2], barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t shTileBytes = (uint32_t)p.BN * 128u;             // one shifted copy [BN i][32 px]
    const uint32_t rawBytes = (uint32_t)p.BN * (kRawW * 4);
    const uint32_t dyRing = base;
    const uint32_t shRing = base + kDySlots * kDyTileBytes;          // 2 stages x 3 copies (1024-aligned: BN % 8 == 0)
    const uint32_t rawRing = shRing + 6u * shTileBytes;              // 2 stages

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
        for (int s = 0; s < 2; s++) {
            mbar_init(smem_u32(&barRawFull[s]), 1); mbar_init(smem_u32(&barRawEmpty[s]), 4);
            mbar_init(smem_u32(&barShFull[s]), 4); mbar_init(smem_u32(&barShEmpty[s]), 1);
        }
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
                const uint32_t s = st & 1;
                if (st >= 2) mbar_wait(smem_u32(&barRawEmpty[s]), ((st >> 1) - 1) & 1);
                const uint32_t full = smem_u32(&barRawFull[s]);
                mbar_expect_tx_elect(full, rawBytes);
                tma_load_4d_elect(rawRing + s * rawBytes, &mapX, full, ox0 - 4, rA + t, i0, n);
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer: nine accumulators, D = F32, A = B = TF32, both K-major, N = BN, M = 128 ----------------
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
#pragma unroll 1
                    for (int ky = 0; ky < 3; ky++) {
                        const uint32_t slot = (gBase + (uint32_t)(t + 2 - ky)) % kDySlots;       // dY row (rA + t) + pad - ky
                        const uint32_t aLo = aLo0 + slot * (kDyTileBytes >> 4);
#pragma unroll
                        for (int kx = 0; kx < 3; kx++)
                            umma_tf32_x4<2, 2>(tmem + (uint32_t)((ky * 3 + kx) * p.BN), aLo, aHi,
                                               bLo0 + (s * 3u + (uint32_t)kx) * shStep, bHi, idesc, first);
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
        // ---------------- shifters: X window [BN i][40 px] -> three copies [BN i][32 px], K-major SWIZZLE_128B ----------------
        unsigned char* gen = smem + (base - smem_u32(smem));
        const float* rawG = reinterpret_cast<const float*>(gen + (rawRing - base));
        unsigned char* shG = gen + (shRing - base);
        const int off0 = 4 - p.pad;                               // window column of chunk pixel 0 for kx = 0
        // 16-byte granule of pixel `lane` inside a 128-byte row, before the swizzle XOR with (row & 7)
        const uint32_t gran = (uint32_t)lane >> 2, sub = ((uint32_t)lane & 3u) << 2;
        uint32_t st = 0;
        for (int cx = 0; cx < p.nCx; cx++) {
            for (int t = 0; t < rows; t++, st++) {
                const uint32_t s = st & 1;
                mbar_wait(smem_u32(&barRawFull[s]), (st >> 1) & 1);
                if (st >= 2) mbar_wait(smem_u32(&barShEmpty[s]), ((st >> 1) - 1) & 1);
                const float* raw = rawG + (size_t)s * (rawBytes >> 2);
                unsigned char* sh = shG + (size_t)s * 3u * shTileBytes;
                for (int i = warp; i < p.BN; i += 4) {
                    const float* rr = raw + i * kRawW + off0 + lane;
                    const uint32_t dst = (uint32_t)i * 128u + (((gran ^ ((uint32_t)i & 7u)) << 4) | sub);
#pragma unroll
                    for (int kx = 0; kx < 3; kx++)
                        *reinterpret_cast<uint32_t*>(sh + kx * shTileBytes + dst) = cvt_tf32(rr[kx]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&barShFull[s]));
                    mbar_arrive(smem_u32(&barRawEmpty[s]));
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
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int i = i0 + c + j;
                        if (c + j < p.BN && i < p.I) atomicAdd(row + i, __uint_as_float(r[j]));
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
    if (N < 1 || I < 1 || O < 1) return SG3_E_INVALID;

    Wg3Params p;
    p.dw = dw; p.N = N; p.I = I; p.O = O; p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.pad = pad; p.ldw = ldw;
    const int nt = (I + 47) / 48;
    p.BN = ((I + nt - 1) / nt + 15) & ~15;                  // 16 / 32 / 48
    p.tilesI = (I + p.BN - 1) / p.BN;
    p.tilesO = (O + 127) / 128;
    p.nCx = (OW + 31) / 32;
    // split the rows so that about two waves of CTAs exist; a split walks at least 8 rows (two extra dY tiles per column chunk)
    const long long baseTiles = (long long)N * p.tilesO * p.tilesI;
    long long splits = ((long long)sg3_sm_count() * 2 + baseTiles - 1) / baseTiles;
    const long long maxSplits = (H + 7) / 8;
    if (splits > maxSplits) splits = maxSplits;
    if (splits < 1) splits = 1;
    p.rowsPerSplit = (int)((H + splits - 1) / splits);
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
    const int smemBytes = kDySlots * kDyTileBytes + 6 * p.BN * 128 + 2 * p.BN * kRawW * 4 + 1024;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([] { return cudaFuncSetAttribute(modconv_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attrErr != cudaSuccess) return (int)attrErr;
    modconv_wgrad3_kernel<<<(unsigned)ctas, kThreadsW3, smemBytes, stream>>>(mapDY, mapX, p);
    return sg3_launch_status();
}
