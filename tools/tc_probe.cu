// Probe for the tcgen05 path: dumps the TMA-written smem tiles and the TMEM accumulator of one tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o tools/tc_probe tools/tc_probe.cu stylegan3-editing_b200/csrc/capi.cu
#include <cstdio>
#include <vector>
#include "../stylegan3-editing_b200/csrc/modconv_tc.cu"

__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, int BN, int tmemCols,
             float* dumpA, float* dumpB, float* dumpD, int variant)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t barFull, barAccum;
    __shared__ uint32_t tmemBase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem) + 1023u) & ~1023u;
    unsigned char* tilesPtr = smem + (tiles - smem_u32(smem));
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&barFull), 1); mbar_init(smem_u32(&barAccum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "r"((uint32_t)tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    const int stageBytes = A_STAGE_BYTES + BN * BK * 4;
    if (threadIdx.x == 128) {
        const uint32_t full = smem_u32(&barFull);
        mbar_expect_tx(full, (uint32_t)stageBytes);
        for (int j = 0; j < 4; j++) tma_load_3d(tiles + j * (BK * 128), &mapX, full, 32 * j, 0, 0);
        tma_load_3d(tiles + A_STAGE_BYTES, &mapW, full, 0, 0, 0);
    }
    mbar_wait(smem_u32(&barFull), 0);
    __syncthreads();
    for (int e = threadIdx.x; e < A_STAGE_BYTES / 4; e += blockDim.x) dumpA[e] = ((float*)tilesPtr)[e];
    for (int e = threadIdx.x; e < BN * BK; e += blockDim.x) dumpB[e] = ((float*)(tilesPtr + A_STAGE_BYTES))[e];
    __syncthreads();
    if (threadIdx.x == 160 && variant == 3) umma_commit(smem_u32(&barAccum));
    if (threadIdx.x == 160 && variant != 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int ks = 0; ks < BK / 8; ks++) {
            uint64_t da, db;
            if (variant == 2) {   // both operands K-major from the W tile: D = W W^T
                const uint32_t idescK = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                da = umma_desc(tiles + A_STAGE_BYTES + ks * 32, 16, 1024); db = da;
                umma_tf32(tmem, da, db, idescK, ks > 0 ? 1u : 0u);
                continue;
            }
            if (variant == 0) { da = umma_desc(tiles + ks * 1024, BK * 128, 512, kLayoutSw128Base32); db = umma_desc(tiles + A_STAGE_BYTES + ks * 32, 16, 1024); }
            else { da = umma_desc(tiles + ks * 1024, 512, BK * 128, kLayoutSw128Base32); db = umma_desc(tiles + A_STAGE_BYTES + ks * 32, 16, 1024); }
            umma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&barAccum));
    }
    if (variant == 3 && warp < 4) {     // TMEM st/ld round trip, no MMA involved
        uint32_t v = 1000u * (32 * warp + lane);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(32 * warp) << 16) + 3u), "r"(__float_as_uint((float)v)) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmemCols) : "memory");
    }
}

int main()
{
    const int I = 32, O = 16, P = 256, BN = 16;
    std::vector<float> hx(I * P), hw(O * I);
    for (int i = 0; i < I; i++) for (int p = 0; p < P; p++) hx[i * P + p] = (float)(i * 1000 + p);
    for (int o = 0; o < O; o++) for (int i = 0; i < I; i++) hw[o * I + i] = (i == (o % 32)) ? 1.f : 0.f;   // y[o][p] = x[o][p]
    float *dx, *dw, *dA, *dB, *dD;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4);
    cudaMalloc(&dA, A_STAGE_BYTES); cudaMalloc(&dB, BN * BK * 4); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap mapX, mapW;
    bool ok1 = make_map3(&mapX, dx, P, I, 1, (uint64_t)P * 4, (uint64_t)P * I * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    bool ok2 = make_map3(&mapW, dw, I, O, 1, (uint64_t)I * 4, (uint64_t)I * O * 4, BK, BN);
    printf("maps %d %d\n", ok1, ok2);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int variant = 0; variant < 4; variant++) {
        cudaMemset(dD, 0xff, 128 * 32 * 4);
        probe_kernel<<<1, 192, 40 * 1024>>>(mapX, mapW, BN, 32, dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        printf("variant %d: sync -> %s\n", variant, cudaGetErrorString(e));
        std::vector<float> hA(A_STAGE_BYTES / 4), hB(BN * BK), hD(128 * 32);
        cudaMemcpy(hA.data(), dA, A_STAGE_BYTES, cudaMemcpyDeviceToHost);
        cudaMemcpy(hB.data(), dB, BN * BK * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hD.data(), dD, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        if (variant == 0) {
            printf("A smem row0 (first 32 floats):"); for (int q = 0; q < 32; q++) printf(" %g", hA[q]); printf("\n");
            printf("A smem row1 (floats 32..63):"); for (int q = 32; q < 64; q++) printf(" %g", hA[q]); printf("\n");
            printf("A box1 row0:"); for (int q = 0; q < 8; q++) printf(" %g", hA[1024 + q]); printf("\n");
            printf("B smem row0:"); for (int q = 0; q < 32; q++) printf(" %g", hB[q]); printf("\n");
            printf("B smem row1:"); for (int q = 32; q < 64; q++) printf(" %g", hB[q]); printf("\n");
        }
        printf("D lane0 cols0..15:"); for (int q = 0; q < 16; q++) printf(" %g", hD[q]); printf("\n");
        printf("D lane1 cols0..15:"); for (int q = 0; q < 16; q++) printf(" %g", hD[32 + q]); printf("\n");
        printf("D lane33 cols0..15:"); for (int q = 0; q < 16; q++) printf(" %g", hD[33 * 32 + q]); printf("\n");
        printf("D lane100 cols0..15:"); for (int q = 0; q < 16; q++) printf(" %g", hD[100 * 32 + q]); printf("\n");
    }
    return 0;
}
