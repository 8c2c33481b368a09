"""filtered_lrelu -- fused bias / FIR-upsample / leaky-ReLU / clamp / FIR-downsample on B200.

Host-side mirror of `torch_utils/ops/filtered_lrelu.py`: same public signature, defaults,
output-size rule, warnings and autograd contract (only the bit-packed sign tensor is saved;
the backward pass is the same op with up/down and filters swapped reading that tensor, so
gradients of any order work).  The native side is `sg3_filtered_lrelu` / `sg3_filtered_lrelu_act`
of libsg3_b200.so instead of the JIT-built `filtered_lrelu_plugin` (filtered_lrelu.cpp:16-298);
the kernels are stateless, so any CUDA stream may be used concurrently (the reference warns,
filtered_lrelu.py:216-217, because of its global filter buffers).
"""
import ctypes
import warnings

import numpy as np
import torch

from . import bias_act as _bias_act
from . import capi
from . import upfirdn2d as _upfirdn2d
from .upfirdn2d import _padding as _parse_padding


def _filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    return int(f.shape[-1]), int(f.shape[0])     # width, height


def _taps_and_shape(f):
    """(host taps or None, width, height-or-0-if-separable) in the C ABI's convention."""
    if f is None:
        return None, 1, 1
    t = _upfirdn2d.host_taps(f)
    if t.ndim == 1:
        return t, t.shape[0], 0
    return t, t.shape[1], t.shape[0]


def output_shape(in_h, in_w, fu, fd, up, down, padding):
    """[out_h, out_w] per filtered_lrelu.py:142-143."""
    fu_w, fu_h = _filter_size(fu)
    fd_w, fd_h = _filter_size(fd)
    px0, px1, py0, py1 = _parse_padding(padding)
    out_w = (in_w * up + (px0 + px1) - (fu_w - 1) - (fd_w - 1) + (down - 1)) // down
    out_h = (in_h * up + (py0 + py1) - (fu_h - 1) - (fd_h - 1) + (down - 1)) // down
    return out_h, out_w


# ---------------------------------------------------------------------------------------------
# Raw launches (no autograd).

def _act_inplace(y, si, sx, sy, gain, slope, clamp, write_signs):
    """gain / lrelu / clamp on `y` in place; returns the sign tensor written (or None)."""
    n, c, h, w = y.shape
    mode, s = capi.SIGNS_NONE, None
    if write_signs:
        s = torch.empty([n, c, h, ((w + 15) & ~15) >> 2], dtype=torch.uint8, device=y.device)
        mode = capi.SIGNS_WRITE
    elif si is not None:
        s, mode = si, capi.SIGNS_READ
        assert s.dtype == torch.uint8 and s.is_contiguous() and s.ndim == 4 and s.shape[:2] == y.shape[:2]
    xs = capi.c_i64x4(*y.stride())
    with torch.cuda.device(y.device):
        rc = capi.lib().sg3_filtered_lrelu_act(
            y.data_ptr(), s.data_ptr() if s is not None else None, n, c, h, w, ctypes.byref(xs),
            s.shape[2] if s is not None else 0, s.shape[3] if s is not None else 0, int(sx), int(sy),
            float(gain), float(slope), float(clamp), mode, capi.dtype_code(y.dtype), capi.stream_ptr(y.device))
    capi.check(rc, 'sg3_filtered_lrelu_act')
    return s if write_signs else None


def _fused(x, fu, fd, b, si, sx, sy, cfg, write_signs, ysum=None, pitched_out=False, round_tf32=False):
    """Try the fused kernel.  Returns (y, signs_written) or None when no specialisation exists.
    `ysum` (float32 [C], zeroed): the kernel adds the per-channel sum of y -- the bias gradient when y = dx.
    `pitched_out`: fp32 outputs whose width is not a multiple of 4 are written with a 16-byte row pitch (returned as a view):
    the gradient wrt a 3x3 conv's output then feeds the tensor-core input-gradient conv without a re-pitching copy.
    `round_tf32`: fp32 outputs are rounded to the nearest TF32 value (see `tf32_rounded_outputs`)."""
    up, down, px0, px1, py0, py1, gain, slope, clamp, flip = cfg
    if x.dtype not in (torch.float16, torch.float32):
        return None
    tu, fuw, fuh = _taps_and_shape(fu)
    td, fdw, fdh = _taps_and_shape(fd)
    L = capi.lib()
    if L.sg3_filtered_lrelu_supported(up, down, fuw, fuh, fdw, fdh) != 0:
        return None
    n, c, ih, iw = x.shape
    oh, ow, sh, swb = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
    rc = L.sg3_filtered_lrelu_shape(ih, iw, up, down, fuw, fuh, fdw, fdh, px0, px1, py0, py1,
                                    ctypes.byref(oh), ctypes.byref(ow), ctypes.byref(sh), ctypes.byref(swb))
    if rc != 0:
        raise RuntimeError('filtered_lrelu: upsampled buffer must be at least the size of the downsampling filter '
                           'and the output at least 1x1')
    if pitched_out and x.dtype == torch.float32 and ow.value % 4 != 0:
        from .modulated_conv import empty_row_pitched
        y = empty_row_pitched([n, c, oh.value, ow.value], x.dtype, x.device)
    else:
        y = torch.empty([n, c, oh.value, ow.value], dtype=x.dtype, device=x.device)
    d = capi.FlreluDesc()
    d.x, d.y = x.data_ptr(), y.data_ptr()
    d.b = b.data_ptr() if b is not None else None
    mode, s = capi.SIGNS_NONE, None
    if write_signs:
        s = torch.empty([n, c, sh.value, swb.value], dtype=torch.uint8, device=x.device)
        mode = capi.SIGNS_WRITE
    elif si is not None:
        s, mode = si, capi.SIGNS_READ
        assert s.dtype == torch.uint8 and s.is_contiguous() and s.ndim == 4 and s.shape[:2] == x.shape[:2]
    d.signs = s.data_ptr() if s is not None else None
    d.fu = tu.ctypes.data if tu is not None else None
    d.fd = td.ctypes.data if td is not None else None
    d.N, d.C, d.inH, d.inW, d.outH, d.outW = n, c, ih, iw, oh.value, ow.value
    es = x.element_size()
    d.xStride = capi.c_i64x4(*[v * es for v in x.stride()])
    d.yStride = capi.c_i64x4(*[v * es for v in y.stride()])
    d.bStride = b.stride(0) * es if b is not None else 0
    d.up, d.down, d.fuW, d.fuH, d.fdW, d.fdH = up, down, fuw, fuh, fdw, fdh
    d.px0, d.py0 = px0, py0
    d.gain, d.slope, d.clamp, d.flip = gain, slope, clamp, int(flip)
    d.signMode = mode
    d.sH, d.sWb = (s.shape[2], s.shape[3]) if s is not None else (0, 0)
    d.sx, d.sy = int(sx), int(sy)
    d.dtype = capi.dtype_code(x.dtype)
    d.flags = capi.FLRELU_ROUND_TF32 if (round_tf32 and x.dtype == torch.float32) else 0
    d.ysum = ysum.data_ptr() if ysum is not None else None
    with torch.cuda.device(x.device):
        rc = L.sg3_filtered_lrelu(ctypes.byref(d), capi.stream_ptr(x.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None
    capi.check(rc, 'sg3_filtered_lrelu')
    return y, (s if write_signs else None)


def _generic(x, fu, fd, b, si, sx, sy, cfg, write_signs):
    """Unfused composition on the same library: bias -> upfirdn2d -> act(+signs) -> upfirdn2d
    (what filtered_lrelu.py:224-230 does when the plugin has no specialised kernel)."""
    up, down, px0, px1, py0, py1, gain, slope, clamp, flip = cfg
    y = x
    if b is not None:
        y = _bias_act._launch(_bias_act._dense(x), b.contiguous(), None, None, None, 0, 1, (1, 0.0, 1.0, -1.0))
    y = _upfirdn2d._run(y, fu, up, up, 1, 1, (px0, px1, py0, py1), flip, up ** 2)
    so = _act_inplace(y, si, sx, sy, gain, slope, clamp, write_signs)
    y = _upfirdn2d._run(y, fd, 1, 1, down, down, (0, 0, 0, 0), flip, 1.0)
    return y, so


class _FilteredLRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fu, fd, b, si, sx, sy, cfg):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        capi.require_cuda(x, 'filtered_lrelu')
        if x.numel() == 0:
            raise RuntimeError('filtered_lrelu: x is empty')
        if b is not None and b.dtype != x.dtype:
            raise TypeError('filtered_lrelu: x and b must have the same dtype')
        write_signs = si is None and (x.requires_grad or (b is not None and b.requires_grad))
        strides = [x.stride(i) for i in range(x.ndim) if x.size(i) > 1]
        if any(a < c for a, c in zip(strides[:-1], strides[1:])):
            warnings.warn('low-performance memory layout detected in filtered_lrelu input', RuntimeWarning)
        res = _fused(x, fu, fd, b, si, sx, sy, cfg, write_signs, round_tf32=(si is None and _rounding_wanted()))
        if res is None:
            if not _quiet_fallback:
                warnings.warn('filtered_lrelu called with parameters that have no fused sm_100a kernel, '
                              'using generic composition', RuntimeWarning)
            res = _generic(x, fu, fd, b, si, sx, sy, cfg, write_signs)
        y, so = res
        ctx.save_for_backward(si if si is not None else so)
        ctx.filters = (fu, fd)
        ctx.cfg = cfg
        ctx.x_hw = (x.shape[2], x.shape[3])
        ctx.s_ofs = (sx, sy)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        (signs,) = ctx.saved_tensors
        fu, fd = ctx.filters
        up, down, px0, px1, py0, py1, gain, slope, clamp, flip = ctx.cfg
        xh, xw = ctx.x_hw
        yh, yw = dy.shape[2], dy.shape[3]
        sx, sy = ctx.s_ofs
        for i in (1, 2, 4, 5, 6):
            assert not ctx.needs_input_grad[i], 'filtered_lrelu: only x and b are differentiable'
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
            fu_w, fu_h = _filter_size(fu)
            fd_w, fd_h = _filter_size(fd)
            # Adjoint: roles of (fu, up) and (fd, down) swap, filters flip, the activation becomes a
            # lookup in the sign tensor (filtered_lrelu.py:254-264).
            adj = (down, up,
                   (fu_w - 1) + (fd_w - 1) - px0, xw * up - yw * down + px0 - (up - 1),
                   (fu_h - 1) + (fd_h - 1) - py0, xh * up - yh * down + py0 - (up - 1),
                   gain * (up ** 2) / (down ** 2), slope, float('inf'), not flip)
            s_ofs = (sx - (fu_w - 1) + px0, sy - (fu_h - 1) + py0)
            if ctx.needs_input_grad[3] and not torch.is_grad_enabled() and dy.dtype in (torch.float16, torch.float32):
                # first-order backward: the backward kernel also accumulates db = sum(dx) per channel (fp32 atomics),
                # which saves the separate reduction pass over dx of filtered_lrelu.py:268
                ysum = torch.zeros([dy.shape[1]], dtype=torch.float32, device=dy.device)
                res = _fused(dy, fd, fu, None, signs, s_ofs[0], s_ofs[1], adj, False, ysum=ysum, pitched_out=True)
                if res is not None:
                    dx, db = res[0], ysum.to(dy.dtype)
            if dx is None:
                dx = _FilteredLRelu.apply(dy, fd, fu, None, signs, s_ofs[0], s_ofs[1], adj)
        if ctx.needs_input_grad[3] and db is None:
            db = dx.sum([0, 2, 3])
        return dx, None, None, db, None, None, None, None


_quiet_fallback = False

# ---------------------------------------------------------------------------------------------
# Outputs that feed a TF32 tensor-core convolution.  The tensor core TRUNCATES its fp32 operands to TF32 (10 mantissa bits);
# cuDNN's TF32 kernels round them to nearest when they load them, which halves the operand error.  The modulated-conv kernels of
# this package read activations straight from HBM by TMA, so the rounding is done where the activation is produced: inside
# `tf32_rounded_outputs(True)` the fused forward kernel rounds every fp32 output to the nearest TF32 value (two integer
# operations per value, ~2 % of the stencil time).  This is the 'round' policy of `modulated_conv.set_tf32_activation_policy`; the
# default policy ('compensate') leaves the activations alone and folds the expected truncation loss into the conv weights.  Under
# 'round', `networks.SynthesisLayer` switches the rounding on for the layers whose consumer is a TF32 conv, and
# `round_for_tf32_convs = True` (set by `sg3_b200.patch_modulated_conv()`) does the same for the reference's own layer code.
# Never applied in the backward pass (gradients pass straight through).

_round_tf32 = False
round_for_tf32_convs = False


class tf32_rounded_outputs:
    """Context manager: `filtered_lrelu` forward outputs (fp32, fused kernel) are rounded to the nearest TF32 value."""

    def __init__(self, enable=True):
        self.enable = bool(enable)

    def __enter__(self):
        global _round_tf32
        self.prev, _round_tf32 = _round_tf32, self.enable
        return self

    def __exit__(self, *exc):
        global _round_tf32
        _round_tf32 = self.prev
        return False


def _rounding_wanted():
    if _round_tf32:
        return True
    if round_for_tf32_convs:
        from .modulated_conv import _math_mode, tf32_activation_policy
        return _math_mode() == 'tf32' and tf32_activation_policy() == 'round'
    return False



def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    """Filtered leaky ReLU for a batch of 2-D images `x` [N, C, H, W] (float16/32, CUDA).

    Per channel: add bias `b`; upsample by `up` (zero insertion), pad/crop by `padding`
    (int | [x, y] | [x0, x1, y0, y1], in upsampled pixels) and convolve with `fu`; multiply by
    `gain`, apply leaky ReLU with `slope`, clamp to +-`clamp`; convolve with `fd` and keep every
    `down`-th pixel.  Filters are float32 [H, W], [taps] (separable) or None (identity).
    Returns [N, C, out_h, out_w] per `output_shape()`.

    impl='cuda' runs the fused kernel (generic composition when no specialisation exists);
    impl='ref' runs the unfused chain bias_act -> upfirdn2d -> bias_act -> upfirdn2d of this same
    library (the reference's PyTorch-op version of that chain lives only in oracle/).
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    px0, px1, py0, py1 = _parse_padding(padding)
    assert gain == float(gain) and gain > 0
    assert slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    for f in (fu, fd):
        assert f is None or (isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2 and f.dtype == torch.float32)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1 and b.shape[0] == x.shape[1]
    capi.require_cuda(x, 'filtered_lrelu')
    if impl == 'ref':
        y = _bias_act.bias_act(x=x, b=b)
        y = _upfirdn2d.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
        y = _bias_act.bias_act(x=y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
        return _upfirdn2d.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)
    cfg = (up, down, px0, px1, py0, py1, float(gain), float(slope),
           float(clamp if clamp is not None else 'inf'), bool(flip_filter))
    return _FilteredLRelu.apply(x, fu, fd, b, None, 0, 0, cfg)
