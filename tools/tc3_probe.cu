// Probe for the 3x3 tensor-core contraction: out-channels on M (A = weights, K-major), pixels on N (B = X, N-major,
// SWIZZLE_128B_ATOM_32B), and the kx shift applied as a COLUMN OFFSET of the TMEM accumulator address.
// Question answered: may tcgen05.mma accumulate into tmem + c for c = 1, 2 (1-column granularity)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o tools/tc3_probe tools/tc3_probe.cu stylegan3-editing_b200/csrc/capi.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../stylegan3-editing_b200/csrc/modconv_tc.cu"

constexpr int NPX = 64;     // pixels per MMA (UMMA N)

__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, float* dumpD, int variant)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barFull, barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t tileW = tiles, tileX = tiles + 128 * BK * 4;       // W: [128 o][32 k] 16 KB, X: NPX/32 boxes of 4 KB
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&barFull), 1); mbar_init(smem_u32(&barAccum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (warp < 4) {      // zero 128 columns of the accumulator
        for (int c = 0; c < 128; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 128) {
        const uint32_t full = smem_u32(&barFull);
        mbar_expect_tx(full, (uint32_t)(128 * BK * 4 + NPX * BK * 4));
        tma_load_3d(tileW, &mapW, full, 0, 0, 0);
        for (int j = 0; j < NPX / 32; j++) tma_load_3d(tileX + j * (BK * 128), &mapX, full, 32 * j, 0, 0);
    }
    mbar_wait(smem_u32(&barFull), 0);
    __syncthreads();
    if (threadIdx.x == 160) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // a K-major (bit15 = 0), b MN-major (bit16 = 1), N = NPX, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(NPX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int nshift = 3;
        for (int sh = 0; sh < nshift; sh++)
            for (int ks = 0; ks < BK / 8; ks++) {
                const uint64_t da = umma_desc(tileW + ks * 32, 16, 1024);
                const uint64_t db = umma_desc(tileX + ks * 1024, BK * 128, 512, kLayoutSw128Base32);
                umma_tf32(tmem + (uint32_t)(variant * sh), da, db, idesc, 1u);
            }
        umma_commit(smem_u32(&barAccum));
    }
    if (warp < 4) {
        mbar_wait(smem_u32(&barAccum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
            for (int j = 0; j < 32; j++) dumpD[(32 * warp + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

int main(int argc, char** argv)
{
    const int vsel = argc > 1 ? atoi(argv[1]) : 4;
    const int I = 32, O = 128, P = 256;
    std::vector<float> hx(I * P), hw(O * I);
    for (int i = 0; i < I; i++) for (int p = 0; p < P; p++) hx[i * P + p] = (float)(i * 256 + p);   // exact in tf32 up to 2047
    for (int o = 0; o < O; o++) for (int i = 0; i < I; i++) hw[o * I + i] = (i == (o % 8)) ? 1.f : 0.f;   // y[o][p] = x[o%8][p]
    float *dx, *dw, *dD;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dD, 128 * 128 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap mapX, mapW;
    bool ok1 = make_map3(&mapX, dx, P, I, 1, (uint64_t)P * 4, (uint64_t)P * I * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    bool ok2 = make_map3(&mapW, dw, I, O, 1, (uint64_t)I * 4, (uint64_t)I * O * 4, BK, 128);
    printf("maps %d %d\n", ok1, ok2);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int variant = vsel; variant <= vsel; variant++) {
        cudaMemset(dD, 0xff, 128 * 128 * 4);
        probe_kernel<<<1, 192, 40 * 1024>>>(mapX, mapW, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        printf("variant %d: sync -> %s\n", variant, cudaGetErrorString(e));
        if (e != cudaSuccess) break;
        std::vector<float> hD(128 * 128);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        const int step = variant, ns = 3;
        int bad = 0;
        for (int o = 0; o < 128; o++) for (int c = 0; c < 128; c++) {
            float ex = 0;
            for (int sh = 0; sh < ns; sh++) { int p = c - step * sh; if (p >= 0 && p < NPX) ex += (float)((o % 8) * 256 + p); }
            if (hD[o * 128 + c] != ex) { if (bad < 6) printf("  mismatch o=%d c=%d got %g want %g\n", o, c, hD[o * 128 + c], ex); bad++; }
        }
        printf("variant %d (%d shifted accumulations, column step %d): %d mismatches\n", variant, ns, step, bad);
        printf("D lane1 cols0..11:"); for (int q = 0; q < 12; q++) printf(" %g", hD[128 + q]); printf("\n");
    }
    return 0;
}
