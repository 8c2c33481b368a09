// Instantiates flrelu_bwd_stream::kernel (dense-up backward shapes): {float, half} x {down 2, down 4} x {no signs, read}.
#include "flrelu_bwd_stream.cuh"

namespace fb = flrelu_bwd_stream;

template <class T, int DOWN, int MODE>
static int launch_one(const fb::Params& p, cudaStream_t stream)
{
    auto kern = fb::kernel<T, DOWN, MODE>;
    const int smem = fb::kWarpsPerCta * fb::Geo<DOWN>::WARP_BYTES;
    static Sg3DeviceOnce once;
    const cudaError_t attrErr = once.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    if (attrErr != cudaSuccess) return (int)attrErr;
    const long long ctas = (p.totalStrips + fb::kWarpsPerCta - 1) / fb::kWarpsPerCta;
    if (ctas > 0x7fffffffLL) return SG3_E_TOOLARGE;
    kern<<<(unsigned)ctas, fb::kWarpsPerCta * 32, smem, stream>>>(p);
    return sg3_launch_status();
}

template <class T, int DOWN> int flrelu_bwd_launch(const fb::Params& p, int signMode, cudaStream_t stream)
{
    if (signMode == SG3_SIGNS_READ) return launch_one<T, DOWN, SG3_SIGNS_READ>(p, stream);
    if (signMode == SG3_SIGNS_NONE) return launch_one<T, DOWN, SG3_SIGNS_NONE>(p, stream);
    return SG3_E_NOKERNEL;
}

template int flrelu_bwd_launch<float, 2>(const fb::Params&, int, cudaStream_t);
template int flrelu_bwd_launch<float, 4>(const fb::Params&, int, cudaStream_t);
template int flrelu_bwd_launch<__half, 2>(const fb::Params&, int, cudaStream_t);
template int flrelu_bwd_launch<__half, 4>(const fb::Params&, int, cudaStream_t);
