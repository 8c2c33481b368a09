// capi.cu -- ABI housekeeping for libsg3_b200.so.
#include "common.cuh"
#include <atomic>

static std::atomic<unsigned long long> g_launches{0};
void sg3_note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
SG3_EXPORT unsigned long long sg3_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

#define SG3_STR2(x) #x
#define SG3_STR(x) SG3_STR2(x)

int sg3_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static std::atomic<int> g_convSmemBudget{0};
int sg3_conv_smem_budget() { return g_convSmemBudget.load(std::memory_order_relaxed); }
SG3_EXPORT int sg3_modconv_set_smem_budget(int bytes)
{
    if (bytes < 0) bytes = 0;
    return g_convSmemBudget.exchange(bytes, std::memory_order_relaxed);
}

#ifdef SG3_TRACE
static std::atomic<unsigned long long*> g_trace{nullptr};
unsigned long long* sg3_trace_buffer() { return g_trace.load(std::memory_order_relaxed); }
SG3_EXPORT void sg3_debug_set_trace(void* buf) { g_trace.store((unsigned long long*)buf, std::memory_order_relaxed); }
#endif

SG3_EXPORT int sg3_abi_version(void) { return SG3_ABI_VERSION; }

SG3_EXPORT int sg3_sizeof_flrelu_desc(void) { return (int)sizeof(sg3_flrelu_desc); }

// SG3_SOURCE_HASH: sha256 (first 12 hex digits) over every file of csrc/ and include/ at build time, passed by build.py, so a
// bench line or a test log names the exact sources the loaded library was compiled from.
#ifndef SG3_SOURCE_HASH
#define SG3_SOURCE_HASH unknown
#endif

SG3_EXPORT const char* sg3_build_info(void)
{
    return "libsg3_b200 sm_100a nvcc " SG3_STR(__CUDACC_VER_MAJOR__) "." SG3_STR(__CUDACC_VER_MINOR__) " abi " SG3_STR(SG3_ABI_VERSION)
           " src " SG3_STR(SG3_SOURCE_HASH);
}

SG3_EXPORT const char* sg3_error_string(int code)
{
    switch (code) {
    case 0: return "success";
    case SG3_E_INVALID: return "invalid arguments";
    case SG3_E_NOKERNEL: return "no fused kernel for these parameters";
    case SG3_E_TOOLARGE: return "size exceeds a limit of this entry point";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}
