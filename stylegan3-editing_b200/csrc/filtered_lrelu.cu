// filtered_lrelu.cu -- host side of the fused filtered leaky-ReLU (geometry, dispatch).
// The fused sm_100a kernel is in flrelu_fused.cuh; this first cut only answers geometry
// queries and reports SG3_E_NOKERNEL so the caller runs the generic composition.
#include "common.cuh"

SG3_EXPORT int sg3_filtered_lrelu_shape(int inH, int inW, int up, int down,
                                        int fuW, int fuH, int fdW, int fdH,
                                        int px0, int px1, int py0, int py1,
                                        int* outH, int* outW, int* sH, int* sWb)
{
    if (inH < 1 || inW < 1 || up < 1 || down < 1 || fuW < 1 || fdW < 1 || fuH < 0 || fdH < 0) return SG3_E_INVALID;
    const int fuh = fuH ? fuH : fuW, fdh = fdH ? fdH : fdW;     // separable filters are square in effect
    const int64_t cw = (int64_t)inW * up + px0 + px1 - (fuW - 1);
    const int64_t ch = (int64_t)inH * up + py0 + py1 - (fuh - 1);
    if (cw <= fdW - 1 || ch <= fdh - 1) return SG3_E_INVALID;  // upsampled buffer smaller than the down filter
    const int64_t yw = (cw - (fdW - 1) + (down - 1)) / down;
    const int64_t yh = (ch - (fdh - 1) + (down - 1)) / down;
    if (yw < 1 || yh < 1 || yw > INT32_MAX || yh > INT32_MAX) return SG3_E_INVALID;
    if (outW) *outW = (int)yw;
    if (outH) *outH = (int)yh;
    const int64_t swActive = yw * down - (down - 1) + (fdW - 1);
    if (sH) *sH = (int)(yh * down - (down - 1) + (fdh - 1));
    if (sWb) *sWb = (int)(((swActive + 15) & ~(int64_t)15) >> 2);
    return 0;
}

SG3_EXPORT int sg3_filtered_lrelu_supported(int up, int down, int fuW, int fuH, int fdW, int fdH)
{
    (void)up; (void)down; (void)fuW; (void)fuH; (void)fdW; (void)fdH;
    return SG3_E_NOKERNEL;
}

SG3_EXPORT int sg3_filtered_lrelu(const sg3_flrelu_desc* d, void* stream)
{
    (void)stream;
    if (!d) return SG3_E_INVALID;
    return SG3_E_NOKERNEL;
}
