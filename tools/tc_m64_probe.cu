// Probe: tcgen05.mma cta_group::1 kind::tf32 with M = 64 -- which TMEM lanes receive the 64 accumulator rows, and may the
// accumulator be placed at a lane offset (16) so that two M = 64 accumulators share the same columns?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o tools/tc_m64_probe tools/tc_m64_probe.cu stylegan3-editing_b200/csrc/capi.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../stylegan3-editing_b200/csrc/modconv_tc.cu"

constexpr int NPX = 32;

__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, float* dumpD, int laneOff)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barFull, barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t tileW = tiles, tileX = tiles + 128 * BK * 4;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&barFull), 1); mbar_init(smem_u32(&barAccum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (warp < 4) {      // fill 64 columns with a marker
        for (int c = 0; c < 64; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c), "r"(__float_as_uint(-7.f)) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 128) {
        const uint32_t full = smem_u32(&barFull);
        mbar_expect_tx(full, (uint32_t)(64 * BK * 4 + NPX * BK * 4));
        tma_load_3d(tileW, &mapW, full, 0, 0, 0);
        tma_load_3d(tileX, &mapX, full, 0, 0, 0);
    }
    mbar_wait(smem_u32(&barFull), 0);
    __syncthreads();
    if (threadIdx.x == 160) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // a K-major, b MN-major, N = NPX, M = 64
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(NPX >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
        for (int ks = 0; ks < BK / 8; ks++) {
            const uint64_t da = umma_desc(tileW + ks * 32, 16, 1024);
            const uint64_t db = umma_desc(tileX + ks * 1024, BK * 128, 512, kLayoutSw128Base32);
            umma_tf32(tmem + ((uint32_t)laneOff << 16), da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&barAccum));
    }
    if (warp < 4) {
        mbar_wait(smem_u32(&barAccum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16), r);
        for (int j = 0; j < 32; j++) dumpD[(32 * warp + lane) * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

int main(int argc, char** argv)
{
    const int laneOff = argc > 1 ? atoi(argv[1]) : 0;
    const int I = 32, O = 64, P = 64;
    std::vector<float> hx(I * P), hw(O * I);
    for (int i = 0; i < I; i++) for (int p = 0; p < P; p++) hx[i * P + p] = (float)(p + 1);            // every channel: pixel index + 1
    for (int o = 0; o < O; o++) for (int i = 0; i < I; i++) hw[o * I + i] = (i == 0) ? (float)(o + 1) : 0.f;   // D[o][p] = (o+1)(p+1)
    float *dx, *dw, *dD;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap mapX, mapW;
    bool ok1 = make_map3(&mapX, dx, P, I, 1, (uint64_t)P * 4, (uint64_t)P * I * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    bool ok2 = make_map3(&mapW, dw, I, O, 1, (uint64_t)I * 4, (uint64_t)I * O * 4, BK, 64);
    printf("maps %d %d laneOff %d\n", ok1, ok2, laneOff);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaMemset(dD, 0, 128 * 32 * 4);
    probe_kernel<<<1, 192, 40 * 1024>>>(mapX, mapW, dD, laneOff);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<float> hD(128 * 32);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    // for every TMEM lane: which accumulator row (value / (p+1) - 1 at column p = 0 and 1) or marker
    for (int l = 0; l < 128; l++) {
        const float v0 = hD[l * 32 + 0], v1 = hD[l * 32 + 1];
        if (v0 == -7.f) printf("lane %3d: untouched\n", l);
        else printf("lane %3d: row %g (col1/2 = %g)\n", l, v0 - 1, v1 / 2 - 1);
    }
    return 0;
}
