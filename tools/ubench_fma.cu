// Microbenchmark: FP32 FMA issue throughput on sm_100a -- scalar FFMA vs packed FFMA2
// (fma.rn.f32x2), register vs kernel-parameter (constant bank) multiplier operands.
// Decides the inner-loop shape of the filtered_lrelu FIR passes.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma ubench_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 8

struct Taps { float t[16]; };

__global__ void __launch_bounds__(256) k_ffma(float* out, Taps tp, float seed)
{
    float acc[NACC * 2];
#pragma unroll
    for (int i = 0; i < NACC * 2; i++) acc[i] = seed * (i + threadIdx.x);
    float v = seed;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC * 2; i++) acc[i] = fmaf(acc[i], tp.t[i & 15], v);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC * 2; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, Taps tp, float seed)
{
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = make_float2(seed * (i + threadIdx.x), seed * i);
    float2 v = make_float2(seed, seed);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = __ffma2_rn(acc[i], make_float2(tp.t[i], tp.t[i]), v);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FIR-like: acc += window[j] * tap[j] with register-resident window pairs and taps from the constant bank.
__global__ void __launch_bounds__(256) k_fir2(float* out, Taps tp, float seed)
{
    float2 w[12];
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = make_float2(seed * (i + threadIdx.x), seed * i);
    float2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS / 6; it++) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int j = 0; j < 12; j++) acc[a] = __ffma2_rn(w[j], make_float2(tp.t[j], tp.t[j]), acc[a]);
#pragma unroll
        for (int i = 0; i < 12; i++) w[i].x += acc[i & 3].y * 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_fir1(float* out, Taps tp, float seed)
{
    float w[24];
#pragma unroll
    for (int i = 0; i < 24; i++) w[i] = seed * (i + threadIdx.x);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int it = 0; it < ITERS / 6; it++) {
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int j = 0; j < 12; j++) acc[a] = fmaf(w[j + (a & 1) * 12], tp.t[j], acc[a]);
#pragma unroll
        for (int i = 0; i < 24; i++) w[i] += acc[i & 7] * 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static void timeit(const char* name, F launch, double fma_per_thread, int blocks, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double fma = fma_per_thread * (double)blocks * threads * reps;
    printf("%-10s %8.3f ms/launch  %8.2f TFMA/s  (%.1f TFLOP/s)  err=%d\n", name, ms / reps, fma / (ms * 1e-3) * 1e-12,
           2 * fma / (ms * 1e-3) * 1e-12, (int)cudaGetLastError());
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs=%d clock=%d kHz\n", sms, clk);
    int blocks = sms * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    Taps tp; for (int i = 0; i < 16; i++) tp.t[i] = 0.999f + 1e-4f * i;
    timeit("ffma",  [&] { k_ffma<<<blocks, threads>>>(out, tp, 1e-3f); },  (double)ITERS * NACC * 2, blocks, threads);
    timeit("ffma2", [&] { k_ffma2<<<blocks, threads>>>(out, tp, 1e-3f); }, (double)ITERS * NACC * 2, blocks, threads);
    timeit("fir1",  [&] { k_fir1<<<blocks, threads>>>(out, tp, 1e-3f); },  (double)(ITERS / 6) * 8 * 12, blocks, threads);
    timeit("fir2",  [&] { k_fir2<<<blocks, threads>>>(out, tp, 1e-3f); },  (double)(ITERS / 6) * 4 * 12 * 2, blocks, threads);
    cudaDeviceSynchronize();
    printf("done err=%d\n", (int)cudaGetLastError());
    return 0;
}
