// modconv_tc.cu -- TF32 tcgen05/TMEM implicit-GEMM contraction for modulated_conv2d (placeholder
// until the tensor-core kernel lands: reports "no kernel" so callers use mathMode 0).
#include "common.cuh"

int sg3_modconv_fwd_tc(const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t)
{
    return SG3_E_NOKERNEL;
}
