// Instantiates flrelu_stream::kernel<__half, 2, *, *, TMA=0> (9 kernels).
#include "flrelu_launch.cuh"

SG3_FLRELU_INSTANTIATE(__half, 2, 0)
