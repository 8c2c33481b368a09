// flrelu_launch.cuh -- launch + mode dispatch of flrelu_stream::kernel for one (dtype, up) pair.
#pragma once

#include "flrelu_stream.cuh"

namespace flrelu_stream {

template <class T, int UP, int FD, int MODE, bool TMA>
int launch_one(const Params& p, cudaStream_t stream)
{
    auto kern = kernel<T, UP, FD, MODE, TMA>;
    const int smem = kWarpsPerCta * Geo<UP>::WARP_BYTES;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([&] {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        // four 28 KB CTAs per SM: ask for the large carve-out
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
    });
    if (attrErr != cudaSuccess) return (int)attrErr;
    const long long ctas = (p.totalStrips + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;
    kern<<<(unsigned)ctas, kWarpsPerCta * 32, smem, stream>>>(p);
    return sg3_launch_status();
}

template <class T, int UP, int FD, bool TMA>
int launch_mode(const Params& p, int mode, cudaStream_t stream)
{
    switch (mode) {
    case SG3_SIGNS_NONE:  return launch_one<T, UP, FD, SG3_SIGNS_NONE, TMA>(p, stream);
    case SG3_SIGNS_WRITE: return launch_one<T, UP, FD, SG3_SIGNS_WRITE, TMA>(p, stream);
    case SG3_SIGNS_READ:  return launch_one<T, UP, FD, SG3_SIGNS_READ, TMA>(p, stream);
    }
    return SG3_E_INVALID;
}

template <class T, int UP, bool TMA>
int launch_fd(const Params& p, int fdMode, int mode, cudaStream_t stream)
{
    switch (fdMode) {
    case 0: return launch_mode<T, UP, 0, TMA>(p, mode, stream);
    case 1: return launch_mode<T, UP, 1, TMA>(p, mode, stream);
    case 2: return launch_mode<T, UP, 2, TMA>(p, mode, stream);
    }
    return SG3_E_INVALID;
}

}  // namespace flrelu_stream

// TMAFLAG: 0 = register-prefetch stage A, 1 = TMA stage A (fp32 only)
#define SG3_FLRELU_INSTANTIATE(T, UP, TMAFLAG)                                                                       \
    template <class TT, int U, int TM> int flrelu_stream_launch(const flrelu_stream::Params&, int, int, cudaStream_t); \
    template <> int flrelu_stream_launch<T, UP, TMAFLAG>(const flrelu_stream::Params& p, int fdMode, int signMode, cudaStream_t stream) \
    { return flrelu_stream::launch_fd<T, UP, (TMAFLAG != 0)>(p, fdMode, signMode, stream); }
