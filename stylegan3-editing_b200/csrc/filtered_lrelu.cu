// filtered_lrelu.cu -- C-ABI entry points of the fused filtered leaky-ReLU: geometry queries,
// specialisation table and launch of the warp-streaming kernel (flrelu_stream.cuh).
//
// Replaces filtered_lrelu.cpp:16-209 of the reference.  Differences by design: taps are passed by
// value in the launch parameters (no setup kernel, no cudaMemcpyToSymbol, no global state -> one launch
// per call and safe on any stream); tiles are scheduled as a 1-D list of warp strips (no grid-z limit,
// 64-bit addressing throughout).
#include <cstring>

#include "flrelu_bwd_stream.cuh"
#include "flrelu_stream.cuh"
#include "tensor_map.h"

namespace fs = flrelu_stream;
namespace fb = flrelu_bwd_stream;

SG3_EXPORT int sg3_filtered_lrelu_shape(int inH, int inW, int up, int down,
                                        int fuW, int fuH, int fdW, int fdH,
                                        int px0, int px1, int py0, int py1,
                                        int* outH, int* outW, int* sH, int* sWb)
{
    if (inH < 1 || inW < 1 || up < 1 || down < 1 || fuW < 1 || fdW < 1 || fuH < 0 || fdH < 0) return SG3_E_INVALID;
    const int fuh = fuH ? fuH : fuW, fdh = fdH ? fdH : fdW;     // a separable filter acts as taps x taps
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

int sg3_flrelu_pointwise(const sg3_flrelu_desc* d, float fuScale, float fdScale, cudaStream_t stream);

static bool is_pointwise(int up, int down, int fuW, int fuH, int fdW, int fdH)
{
    return up == 1 && down == 1 && fuW == 1 && fuH <= 1 && fdW == 1 && fdH <= 1;
}

// Shapes of flrelu_bwd_stream.cuh: up 2 with a dense (<= 12x12) up filter, or up 2 separable with down 4; separable down filter.
static bool is_dense_up_shape(int up, int down, int fuW, int fuH, int fdW, int fdH)
{
    if (up != 2 || (down != 2 && down != 4) || fdH != 0 || fdW < 1 || fdW > 6 * down) return false;
    if (fuW < 1 || fuW > fb::kUpTaps || fuH > fb::kUpTaps) return false;
    return fuH != 0 || down == 4;          // separable up / down 2 belongs to the forward kernel
}

template <class T, int DOWN> int flrelu_bwd_launch(const fb::Params& p, int signMode, cudaStream_t stream);

SG3_EXPORT int sg3_filtered_lrelu_supported(int up, int down, int fuW, int fuH, int fdW, int fdH)
{
    if (is_pointwise(up, down, fuW, fuH, fdW, fdH)) return 0;
    if (is_dense_up_shape(up, down, fuW, fuH, fdW, fdH)) return 0;
    if (down != 2 || (up != 2 && up != 4)) return SG3_E_NOKERNEL;
    if (fuH != 0 || fuW < 1 || fuW > fs::kTapsPerPhase * up) return SG3_E_NOKERNEL;      // separable up filter only
    if (fdW < 1 || fdW > fs::kDownTaps || fdH > fs::kDownTaps) return SG3_E_NOKERNEL;
    return 0;
}

// Kernel instantiations live in flrelu_inst_*.cu (one translation unit per dtype x up factor, built in parallel).
template <class T, int UP, int TMAFLAG> int flrelu_stream_launch(const fs::Params& p, int fdMode, int signMode, cudaStream_t stream);


namespace {

template <class T>
int dispatch_shape(const fs::Params& p, int up, int fdMode, int mode, cudaStream_t stream)
{
    return up == 2 ? flrelu_stream_launch<T, 2, 0>(p, fdMode, mode, stream) : flrelu_stream_launch<T, 4, 0>(p, fdMode, mode, stream);
}

// fp32 input with unit pixel stride and 16-byte aligned base / row / channel / sample strides: stage A by TMA.
bool try_tma_map(fs::Params& p, const sg3_flrelu_desc* d, int up)
{
    if (d->dtype != SG3_F32 || d->xStride[3] != 4) return false;
    if (((uintptr_t)d->x & 15) || (d->xStride[2] & 15) || (d->xStride[1] & 15) || (d->xStride[0] & 15)) return false;
    if (d->xStride[2] <= 0 || d->xStride[1] <= 0 || d->xStride[0] <= 0) return false;
    const uint64_t dims[4] = {(uint64_t)d->inW, (uint64_t)d->inH, (uint64_t)d->C, (uint64_t)d->N};
    const uint64_t strides[3] = {(uint64_t)d->xStride[2], (uint64_t)d->xStride[1], (uint64_t)d->xStride[0]};
    const uint32_t box[4] = {(uint32_t)(up == 2 ? fs::Geo<2>::TIWP : fs::Geo<4>::TIWP), 2, 1, 1};
    return sg3_make_tensor_map(&p.mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace

SG3_EXPORT int sg3_filtered_lrelu(const sg3_flrelu_desc* d, void* stream)
{
    if (!d || !d->x || !d->y) return SG3_E_INVALID;
    if (d->N < 1 || d->C < 1 || d->inH < 1 || d->inW < 1 || d->outH < 1 || d->outW < 1) return SG3_E_INVALID;
    if (d->dtype != SG3_F32 && d->dtype != SG3_F16) return SG3_E_NOKERNEL;
    const int fuW = d->fu ? d->fuW : 1, fuH = d->fu ? d->fuH : 1;
    const int fdW = d->fd ? d->fdW : 1, fdH = d->fd ? d->fdH : 1;
    int rc = sg3_filtered_lrelu_supported(d->up, d->down, fuW, fuH, fdW, fdH);
    if (rc != 0) return rc;
    if (d->signMode != SG3_SIGNS_NONE && (!d->signs || d->sH < 1 || d->sWb < 1)) return SG3_E_INVALID;
    if (d->signMode == SG3_SIGNS_WRITE && (d->sx & 3)) return SG3_E_NOKERNEL;   // sign bytes must align with strips
    // the channel sum of the outputs is produced by the sign-READ (backward) stream kernels only
    if (d->ysum && (d->signMode != SG3_SIGNS_READ || is_pointwise(d->up, d->down, fuW, fuH, fdW, fdH))) return SG3_E_NOKERNEL;
    if (is_pointwise(d->up, d->down, fuW, fuH, fdW, fdH)) {
        // 1x1 "filters" are scalars (a separable 1-tap filter acts on both axes: squared); padding would change the size
        if (d->px0 != 0 || d->py0 != 0 || d->outW != d->inW || d->outH != d->inH) return SG3_E_NOKERNEL;
        const float su = d->fu ? (fuH == 0 ? d->fu[0] * d->fu[0] : d->fu[0]) : 1.0f;
        const float sd = d->fd ? (fdH == 0 ? d->fd[0] * d->fd[0] : d->fd[0]) : 1.0f;
        return sg3_flrelu_pointwise(d, su, sd, (cudaStream_t)stream);
    }
    // The stream kernels fold the gain into the up-filter taps and evaluate lrelu(v) as max(v, slope * v): that needs a
    // positive gain and a slope in [0, 1] (every StyleGAN3 layer: sqrt(2), 0.2).  Anything else takes the generic composition.
    if (d->signMode != SG3_SIGNS_READ && (!(d->gain > 0.f) || !(d->slope >= 0.f && d->slope <= 1.f))) return SG3_E_NOKERNEL;
    if (is_dense_up_shape(d->up, d->down, fuW, fuH, fdW, fdH)) {
        if (d->signMode == SG3_SIGNS_WRITE) return SG3_E_NOKERNEL;
        if ((long long)d->inW * (d->xStride[3] < 0 ? -d->xStride[3] : d->xStride[3]) > 0x7fffffffLL) return SG3_E_NOKERNEL;
        fb::Params q;
        q.x = d->x; q.y = d->y; q.b = d->b; q.s = d->signs; q.ysum = d->ysum;
        q.N = d->N; q.C = d->C; q.inH = d->inH; q.inW = d->inW; q.outH = d->outH; q.outW = d->outW;
        for (int i = 0; i < 4; i++) { q.xs[i] = d->xStride[i]; q.ys[i] = d->yStride[i]; }
        q.bs = d->bStride;
        q.px0 = d->px0; q.py0 = d->py0;
        q.slope = d->slope; q.clamp = d->clamp;
        q.sH = d->sH; q.sWb = d->sWb; q.sx = d->sx; q.sy = d->sy;
        // correlation-ordered dense up taps FU'[a][b] (a separable filter is its outer product), scaled by up^2 * gain
        auto fuAt = [&](int a, int b) -> float {
            const int fh = fuH ? fuH : fuW;
            if (a >= fh || b >= fuW) return 0.f;
            const int sa = d->flip ? a : fh - 1 - a, sb = d->flip ? b : fuW - 1 - b;
            if (!d->fu) return 1.0f;                  // fu == NULL is the 1x1 identity filter (filtered_lrelu.py:160-162 of the reference)
            return fuH ? d->fu[sa * fuW + sb] : d->fu[sa] * d->fu[sb];
        };
        for (int py = 0; py < 2; py++)
            for (int px = 0; px < 2; px++)
                for (int ka = 0; ka < 6; ka++)
                    for (int kb = 0; kb < 6; kb++)
                        q.tu[py][px][ka][kb] = 4.0f * d->gain * fuAt(py + 2 * ka, px + 2 * kb);
        for (int t = 0; t < 24; t++) q.fd[t] = t < fdW ? (d->fd ? d->fd[d->flip ? t : fdW - 1 - t] : 1.0f) : 0.f;
        const int tw = d->down == 2 ? fb::Geo<2>::TW : fb::Geo<4>::TW;
        const long long planes = (long long)d->N * d->C;
        q.stripsX = (d->outW + tw - 1) / tw;
        const long long base = planes * q.stripsX;
        const long long want = (long long)sg3_sm_count() * 16 * 3;
        int chunks = 1;
        if (base < want) {
            chunks = (int)((want + base - 1) / base);
            const int maxChunks = (d->outH + 31) / 32;
            if (chunks > maxChunks) chunks = maxChunks;
            if (chunks < 1) chunks = 1;
        }
        q.chunkRows = (d->outH + chunks - 1) / chunks;
        q.chunksY = (d->outH + q.chunkRows - 1) / q.chunkRows;
        q.totalStrips = base * q.chunksY;
        {
            const long long es = d->dtype == SG3_F32 ? 4 : 2;
            q.vecStore = d->yStride[3] == es && ((uintptr_t)d->y % (2 * es)) == 0 && d->yStride[2] % (2 * es) == 0 &&
                         d->yStride[1] % (2 * es) == 0 && d->yStride[0] % (2 * es) == 0;
        }
        cudaStream_t st = (cudaStream_t)stream;
        if (d->dtype == SG3_F32) return d->down == 2 ? flrelu_bwd_launch<float, 2>(q, d->signMode, st) : flrelu_bwd_launch<float, 4>(q, d->signMode, st);
        return d->down == 2 ? flrelu_bwd_launch<__half, 2>(q, d->signMode, st) : flrelu_bwd_launch<__half, 4>(q, d->signMode, st);
    }

    fs::Params p;
    p.x = d->x; p.y = d->y; p.b = d->b; p.s = d->signs; p.ysum = d->ysum;
    p.N = d->N; p.C = d->C; p.inH = d->inH; p.inW = d->inW; p.outH = d->outH; p.outW = d->outW;
    for (int i = 0; i < 4; i++) { p.xs[i] = d->xStride[i]; p.ys[i] = d->yStride[i]; }
    p.bs = d->bStride;
    p.px0 = d->px0; p.py0 = d->py0;
    p.gain = d->gain; p.slope = d->slope; p.clamp = d->clamp;
    p.sH = d->sH; p.sWb = d->sWb; p.sx = d->sx; p.sy = d->sy;

    // Correlation-ordered taps: F'[b] = f[b] if flip else f[last - b]; zero beyond the real filter.
    const int up = d->up;
    memset(p.tu, 0, sizeof(p.tu));
    memset(p.tv, 0, sizeof(p.tv));
    for (int ph = 0; ph < 4; ph++)
        for (int k = 0; k < fs::kTapsPerPhase; k++) {
            float v = 0.f;
            if (ph < up) {
                const int b = ((up - ph) % up) + up * k;
                if (b < fuW) v = (float)up * (d->fu ? d->fu[d->flip ? b : fuW - 1 - b] : 1.0f);
            }
            p.tu[ph][k] = v;
            p.tv[ph][k] = v * d->gain;
        }
    const bool full = fdH != 0 && d->fd != nullptr;
    float fd2[fs::kDownTaps][fs::kDownTaps];
    for (int b = 0; b < fs::kDownTaps; b++)
        p.fdx[b] = (!full && b < fdW) ? (d->fd ? d->fd[d->flip ? b : fdW - 1 - b] : 1.0f) : 0.f;
    for (int a = 0; a < fs::kDownTaps; a++)
        for (int b = 0; b < fs::kDownTaps; b++)
            fd2[a][b] = (full && a < fdH && b < fdW) ? d->fd[(d->flip ? a : fdH - 1 - a) * fdW + (d->flip ? b : fdW - 1 - b)] : 0.f;
    memset(p.fdr, 0, sizeof(p.fdr));
    for (int rot = 0; rot < 3; rot++)
        for (int i = 0; i < 6; i++) {
            const int k = (i + 2 * rot) % 6;                      // logical accumulator held by slot i at this rotation
            for (int half = 0; half < 2; half++)
                for (int b = 0; b < fs::kDownTaps; b++) p.fdr[rot][half][b >> 1][(b & 1) * 6 + i] = fd2[2 * k + half][b];
        }
    // paired output stores (two adjacent columns per lane): unit pixel stride, every row / plane / base address a multiple
    // of the pair size (strips start at even columns)
    {
        const long long es = d->dtype == SG3_F32 ? 4 : 2;
        p.vecStore = d->yStride[3] == es && ((uintptr_t)d->y % (2 * es)) == 0 && d->yStride[2] % (2 * es) == 0 &&
                     d->yStride[1] % (2 * es) == 0 && d->yStride[0] % (2 * es) == 0;
        if ((d->flags & SG3_FLRELU_ROUND_TF32) && d->dtype == SG3_F32) p.vecStore |= 2;      // bit 1: outputs rounded to TF32
    }

    // Strip decomposition: TW-column strips (58 / 56 outputs for up 2 / 4); rows are chunked only when there are too few strips to
    // fill the machine (one warp per strip, ~16 resident warps per SM, a few waves).
    const long long planes = (long long)d->N * d->C;
    const int tw = up == 2 ? fs::Geo<2>::TW : fs::Geo<4>::TW;
    p.stripsX = (d->outW + tw - 1) / tw;
    const long long base = planes * p.stripsX;
    const long long want = (long long)sg3_sm_count() * 16 * 3;
    int chunks = 1;
    if (base < want) {
        chunks = (int)((want + base - 1) / base);
        const int maxChunks = (d->outH + 31) / 32;            // at least 32 output rows per chunk
        if (chunks > maxChunks) chunks = maxChunks;
        if (chunks < 1) chunks = 1;
    }
    p.chunkRows = (d->outH + chunks - 1) / chunks;
    p.chunksY = (d->outH + p.chunkRows - 1) / p.chunkRows;
    p.totalStrips = base * p.chunksY;
    SG3_TRACE_SET(p);

    // Dense filters that are mirror-symmetric along x (the radial jinc filters are) take the variant that
    // pre-adds mirrored pixels.  Exact tap equality is required; anything else runs the general dense path.
    int fdMode = full ? 1 : 0;
    if (full && fdW == fs::kDownTaps) {
        bool sym = true;
        for (int a = 0; a < fs::kDownTaps && sym; a++)
            for (int b = 0; b < fs::kDownTaps / 2; b++)
                if (fd2[a][b] != fd2[a][fs::kDownTaps - 1 - b]) { sym = false; break; }
        if (sym) fdMode = 2;
    }
    if ((long long)d->inW * (d->xStride[3] < 0 ? -d->xStride[3] : d->xStride[3]) > 0x7fffffffLL) return SG3_E_NOKERNEL;

    cudaStream_t st = (cudaStream_t)stream;
    if (d->dtype == SG3_F32 && try_tma_map(p, d, up))
        return up == 2 ? flrelu_stream_launch<float, 2, 1>(p, fdMode, d->signMode, st)
                       : flrelu_stream_launch<float, 4, 1>(p, fdMode, d->signMode, st);
    if (d->dtype == SG3_F32) return dispatch_shape<float>(p, up, fdMode, d->signMode, st);
    return dispatch_shape<__half>(p, up, fdMode, d->signMode, st);
}
