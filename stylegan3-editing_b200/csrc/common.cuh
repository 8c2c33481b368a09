// common.cuh -- shared helpers for the sm_100a kernels of libsg3_b200.so.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/sg3_b200.h"

#define SG3_EXPORT extern "C" __attribute__((visibility("default")))

// Storage type -> arithmetic type (reference rule, filtered_lrelu.cu:24-45, bias_act.cu:15-18:
// half computes in float and rounds once at the store; double stays double).
template <class T> struct Arith { typedef float type; };
template <> struct Arith<double> { typedef double type; };

template <class T> __device__ __forceinline__ typename Arith<T>::type ld_as(const T* p) { return (typename Arith<T>::type)(*p); }
template <> __device__ __forceinline__ float ld_as<__half>(const __half* p) { return __half2float(*p); }

template <class T> __device__ __forceinline__ void st_as(T* p, typename Arith<T>::type v) { *p = (T)v; }
template <> __device__ __forceinline__ void st_as<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// Every kernel launch of this library is counted (sg3_launch_count) so callers can prove which path ran.
void sg3_note_launches(int n);

static inline int sg3_launch_status(int launches = 1)
{
    sg3_note_launches(launches);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__host__ __device__ __forceinline__ int floor_div(int a, int b)
{
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__host__ __device__ __forceinline__ int pos_mod(int a, int b)
{
    int r = a % b;
    return r < 0 ? r + b : r;
}

// Dynamic shared memory the persistent tensor-core conv kernels may use per CTA (0 = no limit); see sg3_modconv_set_smem_budget.
int sg3_conv_smem_budget();

// Cached SM count of the current device (grid sizing in multiples of the SM count).
int sg3_sm_count();

// One-time per-DEVICE setup of kernel attributes: cudaFuncSetAttribute applies to the current device only, so a
// process that drives several GPUs (G.to('cuda:1'), one thread per device) must repeat it on each of them.
// `f` is idempotent; two threads racing on the same device may both run it.
struct Sg3DeviceOnce {
    std::atomic<int> done[64];
    template <class F> cudaError_t run(F&& f)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return f();
        if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
        const cudaError_t e = f();
        if (e == cudaSuccess) done[dev].store(1, std::memory_order_release);
        return e;
    }
};

// ---- CTA residency trace (debug builds only: -DSG3_TRACE, see tools/overlap_probe.py) ---------------------------------------
// buf[0] = record counter, buf[1] = capacity in records, records of 3 x u64 from buf[2]: (kind << 32 | smid), id, globaltimer.
#ifdef SG3_TRACE
unsigned long long* sg3_trace_buffer();
__device__ __forceinline__ void sg3_trace(unsigned long long* buf, unsigned kind, unsigned long long id)
{
    if (!buf) return;
    unsigned smid;
    unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long slot = atomicAdd(buf, 1ull);
    if (slot < buf[1]) {
        buf[2 + 3 * slot] = ((unsigned long long)kind << 32) | smid;
        buf[3 + 3 * slot] = id;
        buf[4 + 3 * slot] = t;
    }
}
#define SG3_TRACE_FIELD unsigned long long* trace;
#define SG3_TRACE_SET(p) (p).trace = sg3_trace_buffer()
#define SG3_TRACE_EVENT(p, kind, id) sg3_trace((p).trace, kind, id)
#else
#define SG3_TRACE_FIELD
#define SG3_TRACE_SET(p) ((void)0)
#define SG3_TRACE_EVENT(p, kind, id) ((void)0)
#endif
