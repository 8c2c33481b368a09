"""modulated_conv2d -- per-sample modulated convolution with weight demodulation on B200.

Mirror of `modulated_conv2d()` in `models/stylegan3/networks_stylegan3.py:24-63` (same signature
and semantics).  The reference materialises [N,O,I,k,k] weights with ~10 eager kernels and calls a
cuDNN grouped convolution; here the weight chain is one fused prologue kernel
(`sg3_modconv_weights`) and the contraction is `sg3_modconv_fwd`:
  math='fp32'    exact FP32 SIMT contraction (parity mode, <= 1e-5 of the reference's fp32 result)
  math='tf32'    TF32 tcgen05/TMEM implicit GEMM (the reference's default cuDNN path also allows TF32)
  math='fp32x3'  3xTF32 on the tensor cores: both operands split into a TF32 head and tail, three MMAs per K step --
                 fp32-accurate (~1e-6) at tensor-core speed; 1x1 kernels (config R), other shapes run 'fp32' 
The default follows `torch.backends.cudnn.allow_tf32`, like the reference's F.conv2d call.
"""
import torch

from . import capi

_default_math = None      # None -> follow torch.backends.cudnn.allow_tf32

# How the fp32 activations of a TF32 tensor-core conv are treated.  The tensor core truncates them (cuDNN's TF32 kernels round to
# nearest when they load them; ours read them straight from HBM by TMA):
#   'compensate'  (default) the expected truncation loss (a factor 1 - 3.52e-4 on every product) is folded into the modulated weights
#                 by the weight prologue -- free, and measured as accurate as rounding (tiny generators: image error 7.1e-3 / 6.2e-3
#                 with plain truncation, 3.1e-3 / 2.3e-3 compensated, 2.3e-3 / 3.3e-3 rounded; cuDNN: 3.2e-3);
#   'round'       the filtered_lrelu that produces the activations rounds them to the nearest TF32 value
#                 (filtered_lrelu.tf32_rounded_outputs; exact, ~2 % of the stencil time);
#   'truncate'    nothing.
_tf32_policy = 'compensate'
_pitched_output = True    # 3x3 tensor-core convs write a 16-byte row pitch (see conv_forward); False: contiguous outputs


def set_math(mode):
    """Force 'fp32' / 'tf32' for all modulated_conv2d calls, or None to follow cudnn.allow_tf32."""
    global _default_math
    assert mode in (None, 'fp32', 'tf32', 'fp32x3')
    _default_math = mode


def set_tf32_activation_policy(policy):
    """'compensate' | 'round' | 'truncate' (see the comment at `_tf32_policy`)."""
    global _tf32_policy
    assert policy in ('compensate', 'round', 'truncate')
    _tf32_policy = policy


def tf32_activation_policy():
    return _tf32_policy


def _math_mode():
    if _default_math is not None:
        return _default_math
    return 'tf32' if torch.backends.cudnn.allow_tf32 else 'fp32'


def modconv_weights(w, s, demodulate=True, input_gain=None, round_tf32=False, transpose=False, tap_major=False, half=False,
                    split=False, dgrad_taps=False, compensate=False):
    """[N, O, ldw >= I*k*k] float32 modulated (+demodulated, +input-gain) weights, rows zero padded  (:39-56).
    transpose=True (1x1 only): [N, I, ldw >= O], the weight operand of the input-gradient GEMM.
    tap_major=True: [N, k*k, O, ldw >= I], the operand of the 3x3 tensor-core kernel.
    half=True: float16 [N, O, ldw] (what `w.to(x.dtype)` of :61 produces for fp16 layers), operand of the fp16 tensor-core kernel;
    with tap_major=True: float16 [N, k*k, O, ldw >= I], operand of the fp16 3x3 kernel.
    split=True: [N, 2, O, ldw], TF32 head and TF32 tail of every weight, operand of the 3xTF32 kernel (math='fp32x3').
    dgrad_taps=True: [N, k*k, I, ldw >= O], taps flipped and channels transposed: the operand of the k x k input-gradient conv.
    compensate=True (with round_tf32): weights scaled by 1 + 3.52e-4, the expected truncation loss of the activations (see `_tf32_policy`)."""
    capi.require_cuda(w, 'modulated_conv2d')
    O, I, kh, kw = w.shape
    assert kh == kw
    N = s.shape[0]
    w = w.detach().contiguous().float()
    s = s.detach().contiguous().float()
    mode, g = 0, None
    if input_gain is not None:
        g = input_gain.detach().float()
        if g.numel() == 1:
            mode, g = 1, g.reshape(1).contiguous()                 # scalar (the synthesis layers' rsqrt(magnitude_ema))
        elif g.ndim == 1 and g.shape[0] == I:
            mode, g = 2, g.contiguous()                            # per input channel
        else:
            mode, g = 3, g.expand(N, I).contiguous()               # per (sample, input channel), broadcast like :55
    layout = 1 if transpose else (2 if tap_major else (3 if dgrad_taps else 0))
    if dgrad_taps:
        assert not transpose and not tap_major and not half and not split
        ldw = (O + 31) // 32 * 32
        wmod = torch.zeros([N, kh * kw, I, ldw], dtype=torch.float32, device=w.device)     # padding columns must be zero
    elif split:
        assert not transpose and not tap_major and not half
        ldw = (I * kh * kw + 31) // 32 * 32
        wmod = torch.empty([N, 2, O, ldw], dtype=torch.float32, device=w.device)
    elif half and tap_major:                       # operand of the fp16 3x3 tensor-core kernel: [N, k*k, O, ldw >= I] float16
        assert not transpose
        ldw = (I + 63) // 64 * 64
        wmod = torch.empty([N, kh * kw, O, ldw], dtype=torch.float16, device=w.device)
    elif half:
        assert not transpose
        ldw = (I * kh * kw + 63) // 64 * 64
        wmod = torch.empty([N, O, ldw], dtype=torch.float16, device=w.device)
    elif tap_major:
        assert not transpose
        ldw = (I + 31) // 32 * 32
        wmod = torch.empty([N, kh * kw, O, ldw], dtype=torch.float32, device=w.device)
    elif transpose:
        assert kh == 1
        ldw = (O + 31) // 32 * 32
        wmod = torch.zeros([N, I, ldw], dtype=torch.float32, device=w.device)      # padding columns must be zero
    else:
        ldw = (I * kh * kw + 31) // 32 * 32        # row pitch: 128-byte multiple for the TMA-fed tensor-core path
        wmod = torch.empty([N, O, ldw], dtype=torch.float32, device=w.device)
    scratch = torch.empty([1], dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = capi.lib().sg3_modconv_weights(w.data_ptr(), s.data_ptr(), g.data_ptr() if g is not None else None, mode,
                                            wmod.data_ptr(), scratch.data_ptr(), N, I, O, kh, ldw, int(bool(demodulate)),
                                            3 if split else (2 if half else ((4 if compensate else 1) if round_tf32 else 0)), layout, capi.stream_ptr(w.device))
    capi.check(rc, 'sg3_modconv_weights')
    return wmod


def conv_forward(x, wmod, O, k, padding, math):
    """y[n,o] = sum_i wmod[n,o,i] (*) x[n,i] with zero padding; x float32 (or float16 with float16 weights) contiguous.

    The 3x3 tensor-core kernel writes its output with a row pitch rounded up to 16 bytes (returned as a [N, O, OH, OW] view of a
    [N, O, OH, pitch] buffer): the filtered_lrelu that consumes it can then stage its input by TMA although OW * 4 is not a
    16-byte multiple (all 3x3 layers of config T)."""
    N, I, H, W = x.shape
    OH, OW = H + 2 * padding - k + 1, W + 2 * padding - k + 1
    xpitch = 0
    if not x.is_contiguous():           # a row-pitched view [N, I, H, W] of a [N, I, H, pitch] buffer (see _row_pitched)
        assert _row_pitched(x)
        xpitch = x.stride(2)
    pitch = (OW + 3) // 4 * 4 if (math == 'tf32' and k == 3 and x.dtype == torch.float32 and _pitched_output) else OW
    ybuf = torch.empty([N, O, OH, pitch], dtype=x.dtype, device=x.device)
    y = ybuf
    if pitch != OW:
        ybuf[..., OW:].zero_()        # padding columns: defined contents, never written by the kernel, never read by the stencil's TMA
        y = ybuf[..., :OW]
    with torch.cuda.device(x.device):
        ldw = wmod.shape[-1]
        rc = capi.lib().sg3_modconv_fwd_pitched(x.data_ptr(), wmod.data_ptr(), ybuf.data_ptr(), N, I, O, H, W, k, padding, ldw,
                                                xpitch, 0 if pitch == OW else pitch, {'fp32': 0, 'tf32': 1, 'fp32x3': 2}[math],
                                                capi.dtype_code(x.dtype), capi.stream_ptr(x.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None             # e.g. a base pointer off the 16-byte TMA alignment: the caller reruns the SIMT contraction
    capi.check(rc, 'sg3_modconv_fwd')
    return y


def _row_pitched(t):
    """True for a [N, C, H, W] view of a dense [N, C, H, pitch] buffer (unit pixel stride, rows `pitch` apart)."""
    n, c, h, w = t.shape
    p = t.stride(2)
    return t.stride(3) == 1 and p >= w and t.stride(1) == h * p and t.stride(0) == c * h * p


def empty_row_pitched(shape, dtype, device, align=4):
    """[N, C, H, W] tensor whose rows are padded to a multiple of `align` elements (a view when W is not one already)."""
    n, c, h, w = shape
    p = (w + align - 1) // align * align
    buf = torch.empty([n, c, h, p], dtype=dtype, device=device)
    if p == w:
        return buf
    buf[..., w:].zero_()              # the <= 3 padding columns of every row: defined contents (no kernel ever writes them)
    return buf[..., :w]


def tc_supported(I, O, H, W, k, padding):
    """True when sg3_modconv_fwd has a tensor-core kernel for this shape (decides math mode and weight layout)."""
    return capi.lib().sg3_modconv_tc_supported(I, O, H, W, k, padding) == 0


def _reference_formula(x, w, s, demodulate, padding, input_gain):
    """The expression of networks_stylegan3.py:39-62 in torch ops; used only to differentiate."""
    N = x.shape[0]
    O, I, kh, kw = w.shape
    if demodulate:
        w = w * w.square().mean([1, 2, 3], keepdim=True).rsqrt()
        s = s * s.square().mean().rsqrt()
    w = w.unsqueeze(0) * s.unsqueeze(1).unsqueeze(3).unsqueeze(4)
    if demodulate:
        w = w * (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().unsqueeze(2).unsqueeze(3).unsqueeze(4)
    if input_gain is not None:
        w = w * input_gain.expand(N, I).unsqueeze(1).unsqueeze(3).unsqueeze(4)
    y = torch.nn.functional.conv2d(x.reshape(1, -1, *x.shape[2:]), w.reshape(-1, I, kh, kw).to(x.dtype),
                                   padding=padding, groups=N)
    return y.reshape(N, -1, *y.shape[2:])


class _ModConv(torch.autograd.Function):
    """Forward on the sm_100a kernels.  Backward (PTI / fine-tuning): 1x1 kernels with TF32 math run native
    tcgen05 dgrad / wgrad GEMMs (`_native_backward_1x1`); 3x3 kernels, exact-fp32 mode and higher-order
    gradients differentiate the reference expression with library convolutions."""

    @staticmethod
    def forward(ctx, x, w, s, input_gain, demodulate, padding, math):
        O, I, k, _ = w.shape
        xin = x.contiguous()
        if (xin.dtype == torch.float16 and math == 'tf32'
                and ((k == 1 and padding == 0 and (xin.shape[2] * xin.shape[3]) % 8 == 0) or (k == 3 and padding in (0, 2)))):
            # fp16 layer (reference :61 casts the weights to fp16 and runs an fp16 cuDNN conv): fp16 tensor-core kernels,
            # no up / down casts of the activations.  The 3x3 kernel reads x by TMA: rows must be 16 bytes = 8 pixels apart
            # (every config-T width is 4 mod 8, so x goes through one row-pitched copy).
            xk = xin
            if k == 3 and (xin.shape[3] % 8 != 0 or xin.data_ptr() % 16 != 0):
                xk = empty_row_pitched(xin.shape, xin.dtype, xin.device, align=8)
                xk.copy_(xin)
            wmod = modconv_weights(w, s, demodulate=demodulate, input_gain=input_gain, half=True, tap_major=(k == 3))
            y = conv_forward(xk, wmod, O, k, padding, 'tf32')
            if y is not None:
                ctx.save_for_backward(x, w, s, input_gain)
                ctx.cfg = (demodulate, padding, math)
                return y
            math = 'fp32'                  # no fp16 tensor-core launch for this tensor (alignment): upcast + SIMT below
        x32 = xin if xin.dtype == torch.float32 else xin.float()
        if math == 'fp32x3' and (k != 1 or xin.dtype != torch.float32):
            math = 'fp32'                  # the operand-split kernel exists for fp32 1x1 convs
        if math in ('tf32', 'fp32x3') and not tc_supported(I, O, x32.shape[2], x32.shape[3], k, padding):
            math = 'fp32'                  # shapes without a tensor-core kernel run the exact SIMT contraction
        wmod = modconv_weights(w, s, demodulate=demodulate, input_gain=input_gain, round_tf32=(math == 'tf32'),
                               tap_major=(math == 'tf32' and k > 1), split=(math == 'fp32x3'),
                               compensate=(math == 'tf32' and _tf32_policy == 'compensate'))
        y = conv_forward(x32, wmod, O, k, padding, math)
        if y is None and math in ('tf32', 'fp32x3'):
            # the shape has a tensor-core kernel but this tensor does not (unaligned view, tensor-map encode failure):
            # rebuild the weights in the plain fp32 layout and run the exact SIMT contraction
            math = 'fp32'
            wmod = modconv_weights(w, s, demodulate=demodulate, input_gain=input_gain)
            y = conv_forward(x32, wmod, O, k, padding, math)
        if y is None:
            raise capi.Sg3Error('sg3_modconv_fwd: no kernel for these parameters')
        ctx.save_for_backward(x, w, s, input_gain)
        ctx.cfg = (demodulate, padding, math)
        return y if x.dtype == torch.float32 else y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x, w, s, input_gain = ctx.saved_tensors
        demodulate, padding, math = ctx.cfg
        need = ctx.needs_input_grad[:3]
        higher_order = torch.is_grad_enabled()
        if (math == 'tf32' and w.shape[-1] == 1 and padding == 0 and not higher_order
                and x.dtype == torch.float32 and dy.dtype == torch.float32
                and (x.shape[2] * x.shape[3]) % 4 == 0):        # TMA needs 16-byte aligned plane pitches
            return _native_backward_1x1(x, w, s, input_gain, dy, demodulate, need) + (None, None, None, None)
        if (math == 'tf32' and w.shape[-1] == 3 and padding in (0, 2) and not higher_order
                and x.dtype == torch.float32 and dy.dtype == torch.float32):
            res = _native_backward_3x3(x, w, s, input_gain, dy, demodulate, padding, need)
            if res is not None:
                return res + (None, None, None, None)
        with torch.enable_grad(), torch.backends.cudnn.flags(allow_tf32=(math == 'tf32')):
            xs = x.detach().requires_grad_(need[0])
            ws = w.detach().requires_grad_(need[1])
            ss = s.detach().requires_grad_(need[2])
            y = _reference_formula(xs, ws, ss, demodulate, padding, input_gain)
            ins = [t for t, n in zip((xs, ws, ss), need) if n]
            grads = list(torch.autograd.grad(y, ins, dy, create_graph=higher_order)) if ins else []
        out = [grads.pop(0) if n else None for n in need]
        return out[0], out[1], out[2], None, None, None, None


def _gain_arg(input_gain, N, I):
    """(gainMode, contiguous float32 tensor or None) as sg3_modconv_weights expects."""
    if input_gain is None:
        return 0, None
    g = input_gain.detach().float()
    if g.numel() == 1:
        return 1, g.reshape(1).contiguous()
    if g.ndim == 1 and g.shape[0] == I:
        return 2, g.contiguous()
    return 3, g.expand(N, I).contiguous()


def _weights_backward(dWn, w, s, input_gain, demodulate, ldw):
    """Chain rule through the weight prologue (:39-56) on one kernel pair (`sg3_modconv_weights_bwd`): gradient wrt the
    per-sample weights [N, O, ldw] -> (dw [O, I], ds [N, I]).  None when the library has no kernel (I > 2048)."""
    O, I = w.shape[0], w.shape[1]
    N = s.shape[0]
    w2 = w.detach().reshape(O, I).float().contiguous()
    s2 = s.detach().float().contiguous()
    mode, g = _gain_arg(input_gain, N, I)
    dw = torch.empty([O, I], dtype=torch.float32, device=w.device)
    ds = torch.empty([N, I], dtype=torch.float32, device=w.device)
    scratch = torch.empty([1 + N * I], dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = capi.lib().sg3_modconv_weights_bwd(dWn.data_ptr(), w2.data_ptr(), s2.data_ptr(), g.data_ptr() if g is not None else None,
                                                mode, dw.data_ptr(), ds.data_ptr(), scratch.data_ptr(), N, I, O, ldw,
                                                int(bool(demodulate)), capi.stream_ptr(w.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None
    capi.check(rc, 'sg3_modconv_weights_bwd')
    return dw, ds


def _native_backward_1x1(x, w, s, input_gain, dy, demodulate, need):
    """dx, dw, ds of the 1x1 modulated conv on the tcgen05 kernels (TF32): dgrad = the forward GEMM with the
    transposed modulated weights; wgrad = split-K GEMM over pixels; then the chain rule through the weight
    prologue (networks_stylegan3.py:39-56) on the small [N, O, I] tensors."""
    N, I, H, W = x.shape
    O = w.shape[0]
    dy = dy.contiguous()
    xc = x.contiguous()
    dx = dw = ds = None
    if need[0]:
        wT = modconv_weights(w, s, demodulate=demodulate, input_gain=input_gain, round_tf32=True, transpose=True)
        dx = conv_forward(dy, wT, I, 1, 0, 'tf32')
        if dx is None:
            raise capi.Sg3Error('sg3_modconv_fwd (dgrad): no tensor-core kernel for this dy tensor')
    if need[1] or need[2]:
        ldw = (I + 31) // 32 * 32
        dWn = torch.zeros([N, O, ldw], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = capi.lib().sg3_modconv_wgrad(dy.data_ptr(), xc.data_ptr(), dWn.data_ptr(), N, I, O, H, W, ldw,
                                              capi.stream_ptr(x.device))
        capi.check(rc, 'sg3_modconv_wgrad')
        fused = _weights_backward(dWn, w, s, input_gain, demodulate, ldw)
        if fused is not None:
            dw, ds = fused
            return dx, (dw.reshape(w.shape).to(w.dtype) if need[1] else None), (ds.to(s.dtype) if need[2] else None)
        dW = dWn[:, :, :I]                                             # grad wrt the final per-sample weights [N, O, I]
        w2 = w.detach().reshape(O, I).float()
        s2 = s.detach().float()
        if input_gain is not None:
            g = input_gain.detach().float()
            g = g.reshape(1, 1) if g.numel() == 1 else (g.reshape(1, I) if g.ndim == 1 else g.expand(N, I))
            dW = dW * g.unsqueeze(1)
        if demodulate:
            rw = w2.square().mean(dim=1, keepdim=True).rsqrt()          # [O, 1]
            rs = s2.square().mean().rsqrt()
            wn, sn = w2 * rw, s2 * rs
            Wm = wn.unsqueeze(0) * sn.unsqueeze(1)                       # [N, O, I]
            q = Wm.square().sum(dim=2) + 1e-8                            # [N, O]
            d = q.rsqrt()
            dd = (dW * Wm).sum(dim=2)
            dWm = dW * d.unsqueeze(2) - Wm * (dd * q.pow(-1.5)).unsqueeze(2)
            dwn = (dWm * sn.unsqueeze(1)).sum(dim=0)                     # [O, I]
            dsn = (dWm * wn.unsqueeze(0)).sum(dim=1)                     # [N, I]
            dw = rw * dwn - w2 * rw.pow(3) * (dwn * w2).sum(dim=1, keepdim=True) / I
            ds = rs * dsn - s2 * rs.pow(3) * (dsn * s2).sum() / (N * I)
        else:
            dw = (dW * s2.unsqueeze(1)).sum(dim=0)
            ds = (dW * w2.unsqueeze(0)).sum(dim=1)
        dw = dw.reshape(w.shape).to(w.dtype) if need[1] else None
        ds = ds.to(s.dtype) if need[2] else None
    return dx, dw, ds


def _weights_formula(w, s, demodulate, input_gain, N):
    """Per-sample weights [N, O, I, k, k] as a differentiable torch expression (networks_stylegan3.py:39-56)."""
    I = w.shape[1]
    if demodulate:
        w = w * w.square().mean([1, 2, 3], keepdim=True).rsqrt()
        s = s * s.square().mean().rsqrt()
    w = w.unsqueeze(0) * s.unsqueeze(1).unsqueeze(3).unsqueeze(4)
    if demodulate:
        w = w * (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().unsqueeze(2).unsqueeze(3).unsqueeze(4)
    if input_gain is not None:
        w = w * input_gain.expand(N, I).unsqueeze(1).unsqueeze(3).unsqueeze(4)
    return w


def _tma_rows(t):
    """`t` itself when its rows are TMA-addressable (row-pitched view or dense, 16-byte pitch and base), else a row-pitched copy."""
    if _row_pitched(t) and t.stride(2) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    buf = empty_row_pitched(t.shape, t.dtype, t.device)
    buf.copy_(t)
    return buf


def _wgrad3_taps(x, dy, padding):
    """Tap-major per-sample weight gradient [N, 9, O, ldw] (and ldw) from `sg3_modconv_wgrad3`, or None (no kernel for the call)."""
    N, I, H, W = x.shape
    O = dy.shape[1]
    xp, dyp = _tma_rows(x), _tma_rows(dy)
    ldw = (I + 3) // 4 * 4
    dWt = torch.zeros([N, 9, O, ldw], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = capi.lib().sg3_modconv_wgrad3(dyp.data_ptr(), xp.data_ptr(), dWt.data_ptr(), N, I, O, H, W, int(padding), ldw,
                                           dyp.stride(2), xp.stride(2), capi.stream_ptr(x.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None
    capi.check(rc, 'sg3_modconv_wgrad3')
    return dWt, ldw


def conv3x3_weight_grad(x, dy, padding):
    """Per-sample weight gradient of the 3x3 conv on the tcgen05 kernel (`sg3_modconv_wgrad3`, TF32 operands, fp32 accumulation):
    dW[n, o, i, ky, kx] = sum dy[n, o, oy, ox] * x[n, i, oy + ky - pad, ox + kx - pad] -- what the reference gets from the grouped
    `conv2d_weight` of conv2d_gradfix.py:153-174.  x [N, I, H, W], dy [N, O, H + 2 pad - 2, W + 2 pad - 2], float32 (dense or
    row-pitched views; anything else is copied into a row-pitched buffer).  Returns [N, O, I, 3, 3] (a permuted view of the
    tap-major buffer the kernel accumulates into) or None when the library has no kernel for the call."""
    N, I = x.shape[:2]
    O = dy.shape[1]
    res = _wgrad3_taps(x, dy, padding)
    if res is None:
        return None
    return res[0][..., :I].reshape(N, 3, 3, O, I).permute(0, 3, 4, 1, 2)


def _weights_backward_taps(dWt, w, s, input_gain, demodulate, ldw):
    """Chain rule through the weight prologue for k x k kernels on one kernel pair (`sg3_modconv_weights_bwd_taps`): tap-major
    gradient wrt the per-sample weights [N, k*k, O, ldw] -> (dw [O, I, k, k], ds [N, I]).  None when the library has no kernel."""
    O, I, k, _ = w.shape
    N = s.shape[0]
    w2 = w.detach().float().contiguous()
    s2 = s.detach().float().contiguous()
    mode, g = _gain_arg(input_gain, N, I)
    dw = torch.empty([O, I, k, k], dtype=torch.float32, device=w.device)
    ds = torch.empty([N, I], dtype=torch.float32, device=w.device)
    scratch = torch.empty([1 + N * I], dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = capi.lib().sg3_modconv_weights_bwd_taps(dWt.data_ptr(), w2.data_ptr(), s2.data_ptr(), g.data_ptr() if g is not None else None,
                                                     mode, dw.data_ptr(), ds.data_ptr(), scratch.data_ptr(), N, I, O, k, ldw,
                                                     int(bool(demodulate)), capi.stream_ptr(w.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None
    capi.check(rc, 'sg3_modconv_weights_bwd_taps')
    return dw, ds


def _native_backward_3x3(x, w, s, input_gain, dy, demodulate, padding, need):
    """dx of the 3x3 modulated conv on the tcgen05 kernel: the input gradient of a stride-1 convolution is the convolution of
    dy with the flipped, channel-transposed taps and padding 2 - pad -- the SAME implicit-GEMM kernel with the weight prologue
    writing layout 3.  dy arrives from the filtered_lrelu backward kernel as a row-pitched view (its width, e.g. 1046, is not a
    TMA pitch), otherwise it is copied into one.  dw / ds: the per-sample weight gradient is the tcgen05 kernel of
    `sg3_modconv_wgrad3`, then the chain rule through the weight prologue on one kernel pair (`sg3_modconv_weights_bwd_taps`;
    autograd through the weight expression beyond 512 input channels).  None when the kernels have no plan for the shape."""
    N, I, H, W = x.shape
    O = w.shape[0]
    dx = dw = ds = None
    dy_p = None
    if need[0]:
        if not (_row_pitched(dy) and dy.stride(2) % 4 == 0 and dy.data_ptr() % 16 == 0):
            buf = empty_row_pitched(dy.shape, dy.dtype, dy.device)
            buf.copy_(dy)
            dy_p = buf
        else:
            dy_p = dy
        wT = modconv_weights(w, s, demodulate=demodulate, input_gain=input_gain, round_tf32=True, dgrad_taps=True)
        global _pitched_output
        keep, _pitched_output = _pitched_output, False            # dx feeds the previous layer's stencil backward as a plain tensor
        try:
            dx = conv_forward(dy_p, wT, I, 3, 2 - padding, 'tf32')
        finally:
            _pitched_output = keep
        if dx is None:
            return None
    if need[1] or need[2]:
        res = _wgrad3_taps(x, dy if dy_p is None else dy_p, padding)
        if res is None:
            return None
        dWt, ldw = res
        fused = _weights_backward_taps(dWt, w, s, input_gain, demodulate, ldw)
        if fused is not None:
            dw = fused[0].to(w.dtype) if need[1] else None
            ds = fused[1].to(s.dtype) if need[2] else None
            return dx, dw, ds
        dW = dWt[..., :I].reshape(N, 3, 3, O, I).permute(0, 3, 4, 1, 2)          # more than 512 input channels: autograd through the expression
        with torch.enable_grad():
            ws = w.detach().requires_grad_(need[1])
            ss = s.detach().requires_grad_(need[2])
            Wn = _weights_formula(ws, ss, demodulate, input_gain, N)
            ins = [t for t, n in zip((ws, ss), need[1:]) if n]
            grads = list(torch.autograd.grad(Wn, ins, dW))
        dw = grads.pop(0) if need[1] else None
        ds = grads.pop(0) if need[2] else None
    return dx, dw, ds


def modulated_conv2d(x, w, s, demodulate=True, padding=0, input_gain=None, math=None):
    """x [N, I, H, W], w [O, I, k, k], s [N, I] -> [N, O, H+2p-k+1, W+2p-k+1].

    w is pre-normalised per output channel and s by the batch RMS when `demodulate`; the modulated
    weights are demodulated per (sample, output channel); `input_gain` ([], [I] or [N, I]) scales
    input channels.  `math` overrides the module default ('fp32' | 'tf32').
    """
    N = int(x.shape[0])
    O, I, kh, kw = w.shape
    assert x.ndim == 4 and x.shape[1] == I and tuple(s.shape) == (N, I) and kh == kw
    if isinstance(padding, (list, tuple)):
        assert padding[0] == padding[1]
        padding = padding[0]
    capi.require_cuda(x, 'modulated_conv2d')
    return _ModConv.apply(x, w, s, input_gain, bool(demodulate), int(padding), math or _math_mode())
