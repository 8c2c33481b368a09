// Issue-rate / operand-fetch microbenchmark for tcgen05.mma kind::tf32 with smem operands (no TMA, operands resident):
// how many clocks does one M=128 x N x K=8 MMA take for the operand layouts used by modconv_tc.cu / modconv_tc3.cu?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o tools/tc_mma_bench tools/tc_mma_bench.cu stylegan3-editing_b200/csrc/capi.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../stylegan3-editing_b200/csrc/modconv_tc.cu"

// variant 0: A K-major SW128, B MN-major SW128_BASE32   (3x3 kernel)
// variant 1: A MN-major SW128_BASE32, B K-major SW128   (1x1 kernel)
// variant 2: both K-major                                (wgrad kernel)
__global__ void __launch_bounds__(192, 1)
bench_kernel(int variant, int N, int iters, int unrollSame, long long* out, int dmode)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;
    for (int e = threadIdx.x; e < 96 * 1024 / 4; e += blockDim.x) ((float*)(smem + (tiles - smem_u32(smem))))[e] = 0.f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (warp == 5 && dmode >= 20) {
        // whole warp runs the loop (uniform control flow), one elected lane issues
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t aBase = tiles, bBase = tiles + 32 * 1024;
        const uint64_t da0 = umma_desc(aBase, 16, 1024), db0 = umma_desc(bBase, 4096, 512, kLayoutSw128Base32);
        const uint32_t aLo = (uint32_t)da0, aHi = (uint32_t)(da0 >> 32), bLo = (uint32_t)db0, bHi = (uint32_t)(db0 >> 32);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const uint32_t d = dmode == 21 ? tmem + (uint32_t)((it & 3) * 128 + ((it & 1) ? 2 : 0)) : tmem;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
                "elect.sync _|q, 0xffffffff;\n\t"
                "setp.ne.b32 p, %6, 0;\n\t"
                "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 64;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                "add.u32 al, %1, 4;\n\tadd.u32 bl, %3, 128;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                "add.u32 al, %1, 6;\n\tadd.u32 bl, %3, 192;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                ::"r"(d), "r"(aLo), "r"(aHi), "r"(bLo), "r"(bHi), "r"(idesc), "r"(1u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    if (threadIdx.x == 160 && dmode < 20) {
        const uint32_t aMaj = variant == 1 ? 1u : 0u, bMaj = variant == 0 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (aMaj << 15) | (bMaj << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t aBase = tiles, bBase = tiles + 32 * 1024;
        if (dmode >= 10) {
            // lean issue: descriptor = constant high word + (constant | addr >> 4) low word; four MMAs per asm block
            const uint64_t da0 = umma_desc(aBase, 16, 1024), db0 = umma_desc(bBase, 4096, 512, kLayoutSw128Base32);
            const uint32_t aLo = (uint32_t)da0, aHi = (uint32_t)(da0 >> 32), bLo = (uint32_t)db0, bHi = (uint32_t)(db0 >> 32);
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                const uint32_t d = dmode == 11 ? tmem + (uint32_t)((it & 3) * 128 + ((it & 1) ? 2 : 0)) : tmem;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
                    "setp.ne.b32 p, %6, 0;\n\t"
                    "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                    "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 64;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                    "add.u32 al, %1, 4;\n\tadd.u32 bl, %3, 128;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
                    "add.u32 al, %1, 6;\n\tadd.u32 bl, %3, 192;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                    ::"r"(d), "r"(aLo), "r"(aHi), "r"(bLo), "r"(bHi), "r"(idesc), "r"(1u) : "memory");
            }
            const long long t1 = clock64();
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        } else {
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const int k2 = unrollSame ? 0 : ks;
                uint64_t da, db;
                if (variant == 0) { da = umma_desc(aBase + k2 * 32, 16, 1024); db = umma_desc(bBase + k2 * 1024, 4096, 512, kLayoutSw128Base32); }
                else if (variant == 1) { da = umma_desc(aBase + k2 * 1024, 4096, 512, kLayoutSw128Base32); db = umma_desc(bBase + k2 * 32, 16, 1024); }
                else { da = umma_desc(aBase + k2 * 32, 16, 1024); db = umma_desc(bBase + k2 * 32, 16, 1024); }
                // dmode 0: one accumulator; 1: column offset +2; 2: switch among 4 accumulators every 4 MMAs; 3: switch every MMA;
                // 4: like 2 but one of them at offset +2
                uint32_t d = tmem;
                if (dmode == 1) d += 2;
                else if (dmode == 2) d += (uint32_t)((it & 3) * 128);
                else if (dmode == 3) d += (uint32_t)(ks * 128);
                else if (dmode == 4) d += (uint32_t)((it & 3) * 128 + ((it & 3) == 1 ? 2 : 0));
                umma_tf32(d, da, db, idesc, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main()
{
    long long* dout; cudaMalloc(&dout, 16);
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    for (int variant = 0; variant < 1; variant++)
        for (int N : {32, 64, 96, 128})
            for (int same : {0, 10, 11, 20, 21}) {
                if ((same == 11 || same == 21) && N > 96) continue;      // four accumulators at column offsets: 4 x N + 2 must fit 512 TMEM columns
                bench_kernel<<<1, 192, 98 * 1024>>>(variant, N, iters, 0, dout, same);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
                printf("variant %d N=%3d dmode=%d: %s  issue %.1f clk/MMA, complete %.1f clk/MMA (floor %d)\n", variant, N, same,
                       cudaGetErrorString(e), (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), N / 2);
            }
    // small-N sweep for the tensor-core FIR question (tools/tc_fir_model.py): banded-Toeplitz FIR formulations need
    // N = 16 .. 64 outputs per MMA; is an M = 128 x N x 8 MMA with smem operands paced by N / 2 clocks or by its operand reads?
    for (int N : {16, 32, 48, 64, 128, 256}) {
        bench_kernel<<<1, 192, 98 * 1024>>>(0, N, iters, 0, dout, 20);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        printf("FIRSWEEP N=%3d: %s issue %.2f complete %.2f clk/MMA  (tensor floor %d, smem operand bytes %d)\n", N, cudaGetErrorString(e),
               (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), N / 2, 128 * 32 + N * 32);
    }
    return 0;
}
