// Instantiates flrelu_stream::kernel<float, 2, *, *, TMA=0> (9 kernels).
#include "flrelu_launch.cuh"

SG3_FLRELU_INSTANTIATE(float, 2, 0)
