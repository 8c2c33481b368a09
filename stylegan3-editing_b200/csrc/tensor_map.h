// tensor_map.h -- host-side TMA tensor-map encoder shared by the TMA-fed kernels (defined in modconv_tc.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

// Tiled tensor map of rank 1..5 through cuTensorMapEncodeTiled (resolved with cudaGetDriverEntryPoint: no link-time libcuda
// dependency).  dims / box in elements, stridesBytes for dimensions 1..rank-1.  Out-of-bounds elements read as zero.
// Returns false when the driver entry point is missing or the encoder rejects the layout (the caller then reports NOKERNEL).
bool sg3_make_tensor_map(CUtensorMap* m, CUtensorMapDataType type, int rank, const void* base, const uint64_t* dims,
                         const uint64_t* stridesBytes, const uint32_t* box, CUtensorMapSwizzle swizzle);
