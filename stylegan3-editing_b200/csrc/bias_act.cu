// bias_act.cu -- y = clamp(act(x + b) * gain) and its first/second derivative forms.
//
// Behaviour follows the reference op (torch_utils/ops/bias_act.py:92-121 for the value,
// bias_act.cu:56-142 for the gradient forms expressed through the saved x / y), written
// here as value / d1 / d2 functors over a 128-bit vectorised, grid-stride elementwise
// kernel: the op is pure HBM streaming (2 * numel * esize bytes), so each thread moves
// 16 bytes per tensor per step and the grid is a multiple of the SM count.
#include "common.cuh"

namespace {

struct BiasActParams {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    int64_t sizeX; int32_t sizeB; int64_t stepB;
    int grad; float alpha, gain, clamp;
    int biasMode;      // 0: none; 1: one bias per 16-byte packet, 32-bit index math; 2: per element, 32-bit; 3: per element, 64-bit
};

template <class S> struct ActConst {
    static __device__ __forceinline__ S selu_scale() { return (S)1.0507009873554804934193349852946; }
    static __device__ __forceinline__ S selu_alpha() { return (S)1.6732632423543772848170429916717; }
};

// value(v), d1(xr, yy), d2(xr, yy): yy = saved output / gain, xr = saved input + bias.
template <int A, class S> struct Act;

template <class S> struct Act<1, S> {  // linear
    static __device__ __forceinline__ S f(S v, S) { return v; }
    static __device__ __forceinline__ S d1(S, S, S) { return (S)1; }
    static __device__ __forceinline__ S d2(S, S, S) { return (S)0; }
};
template <class S> struct Act<2, S> {  // relu
    static __device__ __forceinline__ S f(S v, S) { return v > 0 ? v : (S)0; }
    static __device__ __forceinline__ S d1(S, S yy, S) { return yy > 0 ? (S)1 : (S)0; }
    static __device__ __forceinline__ S d2(S, S, S) { return (S)0; }
};
template <class S> struct Act<3, S> {  // lrelu
    static __device__ __forceinline__ S f(S v, S a) { return v > 0 ? v : v * a; }
    static __device__ __forceinline__ S d1(S, S yy, S a) { return yy > 0 ? (S)1 : a; }
    static __device__ __forceinline__ S d2(S, S, S) { return (S)0; }
};
template <class S> struct Act<4, S> {  // tanh
    static __device__ __forceinline__ S f(S v, S) { return tanh(v); }
    static __device__ __forceinline__ S d1(S, S yy, S) { return (S)1 - yy * yy; }
    static __device__ __forceinline__ S d2(S, S yy, S) { return ((S)1 - yy * yy) * ((S)-2 * yy); }
};
template <class S> struct Act<5, S> {  // sigmoid
    static __device__ __forceinline__ S f(S v, S) { return (S)1 / ((S)1 + exp(-v)); }
    static __device__ __forceinline__ S d1(S, S yy, S) { return yy * ((S)1 - yy); }
    static __device__ __forceinline__ S d2(S, S yy, S) { return yy * ((S)1 - yy) * ((S)1 - (S)2 * yy); }
};
template <class S> struct Act<6, S> {  // elu
    static __device__ __forceinline__ S f(S v, S) { return v >= 0 ? v : expm1(v); }
    static __device__ __forceinline__ S d1(S, S yy, S) { return yy >= 0 ? (S)1 : yy + (S)1; }
    static __device__ __forceinline__ S d2(S, S yy, S) { return yy >= 0 ? (S)0 : yy + (S)1; }
};
template <class S> struct Act<7, S> {  // selu
    static __device__ __forceinline__ S f(S v, S) {
        return v >= 0 ? ActConst<S>::selu_scale() * v : ActConst<S>::selu_scale() * ActConst<S>::selu_alpha() * expm1(v);
    }
    static __device__ __forceinline__ S d1(S, S yy, S) {
        return yy >= 0 ? ActConst<S>::selu_scale() : yy + ActConst<S>::selu_scale() * ActConst<S>::selu_alpha();
    }
    static __device__ __forceinline__ S d2(S, S yy, S) {
        return yy >= 0 ? (S)0 : yy + ActConst<S>::selu_scale() * ActConst<S>::selu_alpha();
    }
};
template <class S> struct Act<8, S> {  // softplus
    static __device__ __forceinline__ S f(S v, S) { return v > (S)80 ? v : log1p(exp(v)); }
    static __device__ __forceinline__ S d1(S, S yy, S) { return (S)1 - exp(-yy); }
    static __device__ __forceinline__ S d2(S, S yy, S) { S c = exp(-yy); return c * ((S)1 - c); }
};
template <class S> struct Act<9, S> {  // swish; derivatives through the saved input xr
    static __device__ __forceinline__ S f(S v, S) { return v < (S)-80 ? (S)0 : v / ((S)1 + exp(-v)); }
    static __device__ __forceinline__ S d1(S xr, S, S) {
        if (xr > (S)40) return (S)1;
        S c = exp(xr), d = c + (S)1;
        return c * (xr + d) / (d * d);
    }
    static __device__ __forceinline__ S d2(S xr, S, S) {
        if (xr > (S)40) return (S)0;
        S c = exp(xr), d = c + (S)1;
        return c * (xr * ((S)2 - d) + (S)2 * d) / (d * d * d);
    }
};

template <int A, class S>
__device__ __forceinline__ S bias_act_one(S x, S b, S xref, S yref, S dy, int grad, S alpha, S gain, S clamp)
{
    S y;
    if (grad == 0) {
        y = Act<A, S>::f(x + b, alpha) * gain;
        // same results as (y > -c && y < c) ? y : (y >= 0 ? c : -c), including NaN -> -c, in two min/max instructions
        if (clamp >= 0) y = min(max(y, -clamp), clamp);
        return y;
    }
    S xr = xref + b;
    if (A == 9) yref = Act<9, S>::f(xr, alpha) * gain;      // swish saves x, not y (bias_act.py:31)
    S yy = gain != 0 ? yref / gain : (S)0;
    S d = grad == 1 ? Act<A, S>::d1(xr, yy, alpha) : Act<A, S>::d2(xr, yy, alpha);
    y = x * d * gain * dy;
    if (clamp >= 0) y = (yref > -clamp && yref < clamp) ? y : (S)0;
    return y;
}

template <class T> struct Vec;   // 16-byte packet
template <> struct Vec<float>  { enum { N = 4 }; typedef float4 type; };
template <> struct Vec<__half> { enum { N = 8 }; typedef uint4 type; };
template <> struct Vec<double> { enum { N = 2 }; typedef double2 type; };

template <class T, int A, bool VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActParams p)
{
    typedef typename Arith<T>::type S;
    constexpr int VN = VEC ? (int)Vec<T>::N : 1;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b;
    const T* xref = (const T*)p.xref; const T* yref = (const T*)p.yref; const T* dy = (const T*)p.dy;
    T* y = (T*)p.y;
    const int64_t nvec = p.sizeX / VN;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int64_t base = v * VN;
        T xs[VN], xr[VN], yr[VN], ds[VN], out[VN];
        if (VEC) {
            typedef typename Vec<T>::type V;
            *(V*)xs = __ldcs((const V*)(x + base));
            if (xref) *(V*)xr = __ldcs((const V*)(xref + base));
            if (yref) *(V*)yr = __ldcs((const V*)(yref + base));
            if (dy)   *(V*)ds = __ldcs((const V*)(dy + base));
        } else {
            xs[0] = x[base];
            if (xref) xr[0] = xref[base];
            if (yref) yr[0] = yref[base];
            if (dy)   ds[0] = dy[base];
        }
        // Bias index (i / stepB) % sizeB: the 64-bit divisions of the general form cost more than the memory traffic of
        // the whole packet, so the common layouts take a cheaper route (NCHW with H*W a multiple of the packet: one
        // 32-bit division per packet).
        S bvec = (S)0;
        if (p.biasMode == 1) bvec = ld_as<T>(b + ((uint32_t)base / (uint32_t)p.stepB) % (uint32_t)p.sizeB);
        if (p.grad == 0 && p.biasMode <= 1) {
            // forward value with one bias per packet (every NCHW activation): no per-element branches
#pragma unroll
            for (int j = 0; j < VN; j++) {
                S r = Act<A, S>::f(ld_as<T>(xs + j) + bvec, alpha) * gain;
                if (clamp >= 0) r = min(max(r, -clamp), clamp);
                st_as<T>(out + j, r);
            }
        } else
#pragma unroll
        for (int j = 0; j < VN; j++) {
            S bv = bvec;
            if (p.biasMode == 2) bv = ld_as<T>(b + (((uint32_t)base + j) / (uint32_t)p.stepB) % (uint32_t)p.sizeB);
            else if (p.biasMode == 3) bv = ld_as<T>(b + ((base + j) / p.stepB) % p.sizeB);
            S r = bias_act_one<A, S>(ld_as<T>(xs + j), bv, xref ? ld_as<T>(xr + j) : (S)0, yref ? ld_as<T>(yr + j) : (S)0,
                                     dy ? ld_as<T>(ds + j) : (S)1, p.grad, alpha, gain, clamp);
            st_as<T>(out + j, r);
        }
        if (VEC) {
            typedef typename Vec<T>::type V;
            __stcs((V*)(y + base), *(V*)out);
        } else {
            y[base] = out[0];
        }
    }
}

template <class T, int A>
int launch_bias_act(const BiasActParams& p, cudaStream_t stream)
{
    const int VN = (int)Vec<T>::N;
    auto aligned = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
    bool vec = (p.sizeX % VN == 0) && aligned(p.x) && aligned(p.xref) && aligned(p.yref) && aligned(p.dy) && aligned(p.y);
    const int64_t nvec = vec ? p.sizeX / VN : p.sizeX;
    const int threads = 256;
    int64_t blocks = ceil_div64(nvec, threads);
    const int64_t cap = (int64_t)sg3_sm_count() * 16;    // 16 CTAs of 256 threads per SM, grid-stride beyond that
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    BiasActParams q = p;
    const bool small = p.sizeX <= 0xffffffffLL && p.stepB <= 0xffffffffLL;
    if (!p.b) q.biasMode = 0;
    else if (small && vec && p.stepB % VN == 0) q.biasMode = 1;
    else q.biasMode = small ? 2 : 3;
    if (vec) bias_act_kernel<T, A, true><<<(unsigned)blocks, threads, 0, stream>>>(q);
    else     bias_act_kernel<T, A, false><<<(unsigned)blocks, threads, 0, stream>>>(q);
    return sg3_launch_status();
}

template <class T>
int dispatch_act(const BiasActParams& p, int act, cudaStream_t stream)
{
    switch (act) {
    case 1: return launch_bias_act<T, 1>(p, stream);
    case 2: return launch_bias_act<T, 2>(p, stream);
    case 3: return launch_bias_act<T, 3>(p, stream);
    case 4: return launch_bias_act<T, 4>(p, stream);
    case 5: return launch_bias_act<T, 5>(p, stream);
    case 6: return launch_bias_act<T, 6>(p, stream);
    case 7: return launch_bias_act<T, 7>(p, stream);
    case 8: return launch_bias_act<T, 8>(p, stream);
    case 9: return launch_bias_act<T, 9>(p, stream);
    }
    return SG3_E_INVALID;
}

}  // namespace

SG3_EXPORT int sg3_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                            int64_t sizeX, int32_t sizeB, int64_t stepB,
                            int grad, int act, float alpha, float gain, float clamp,
                            int dtype, void* stream)
{
    if (!x || !y || sizeX < 0 || grad < 0 || grad > 2 || act < 1 || act > 9) return SG3_E_INVALID;
    if (b && (sizeB <= 0 || stepB <= 0)) return SG3_E_INVALID;
    if (sizeX == 0) return 0;
    BiasActParams p;
    p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
    p.sizeX = sizeX; p.sizeB = b ? sizeB : 1; p.stepB = b ? stepB : 1;
    p.grad = grad; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case SG3_F32: return dispatch_act<float>(p, act, st);
    case SG3_F16: return dispatch_act<__half>(p, act, st);
    case SG3_F64: return dispatch_act<double>(p, act, st);
    }
    return SG3_E_INVALID;
}
