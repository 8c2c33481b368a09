"""bias_act -- fused bias + activation + gain + clamp on B200.

Host-side mirror of `torch_utils/ops/bias_act.py` (same `bias_act()` signature, defaults,
`activation_funcs` table and first/second-order autograd), backed by `sg3_bias_act` of
libsg3_b200.so instead of the JIT-built `bias_act_plugin` (bias_act.cpp:32-90).
"""
import numpy as np
import torch

from . import capi


class _Spec(dict):
    __getattr__ = dict.__getitem__


def _spec(func, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
    return _Spec(func=func, def_alpha=def_alpha, def_gain=def_gain, cuda_idx=cuda_idx, ref=ref, has_2nd_grad=has_2nd_grad)


_F = torch.nn.functional
# name -> evaluation lambda (documentation / oracle cross-checks only), default alpha and gain,
# kernel index, which forward tensor the derivative is expressed through, 2nd-derivative flag.
activation_funcs = {
    'linear':   _spec(lambda x, **_: x,                          0,   1,          1, '',  False),
    'relu':     _spec(lambda x, **_: _F.relu(x),                 0,   np.sqrt(2), 2, 'y', False),
    'lrelu':    _spec(lambda x, alpha, **_: _F.leaky_relu(x, alpha), 0.2, np.sqrt(2), 3, 'y', False),
    'tanh':     _spec(lambda x, **_: torch.tanh(x),              0,   1,          4, 'y', True),
    'sigmoid':  _spec(lambda x, **_: torch.sigmoid(x),           0,   1,          5, 'y', True),
    'elu':      _spec(lambda x, **_: _F.elu(x),                  0,   1,          6, 'y', True),
    'selu':     _spec(lambda x, **_: _F.selu(x),                 0,   1,          7, 'y', True),
    'softplus': _spec(lambda x, **_: _F.softplus(x),             0,   1,          8, 'y', True),
    'swish':    _spec(lambda x, **_: torch.sigmoid(x) * x,       0,   np.sqrt(2), 9, 'x', True),
}


def _dense_like(t, ref):
    """t with the same dense layout as ref (contiguous or channels_last)."""
    if ref.ndim > 2 and ref.stride(1) == 1 and ref.is_contiguous(memory_format=torch.channels_last if ref.ndim == 4 else torch.contiguous_format):
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


def _launch(x, b, xref, yref, dy, grad, dim, cfg):
    """y = kernel(x, ...) for dense x; b indexes dimension `dim`."""
    act_idx, alpha, gain, clamp = cfg
    capi.require_cuda(x, 'bias_act')
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    step_b = x.stride(dim) if b is not None else 1
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(x.device):
        rc = capi.lib().sg3_bias_act(ptr(x), ptr(b), ptr(xref), ptr(yref), ptr(dy), ptr(y),
                                     x.numel(), b.numel() if b is not None else 0, step_b,
                                     grad, act_idx, alpha, gain, clamp,
                                     capi.dtype_code(x.dtype), capi.stream_ptr(x.device))
    capi.check(rc, 'sg3_bias_act')
    return y


def _dense(x):
    if x.ndim == 4 and x.stride(1) == 1 and x.is_contiguous(memory_format=torch.channels_last):
        return x
    return x.contiguous()


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, dim, act, cfg):
        spec = activation_funcs[act]
        x = _dense(x)
        b = b.contiguous() if b is not None else None
        _, _, gain, clamp = cfg
        identity = act == 'linear' and gain == 1 and clamp < 0 and b is None
        y = x if identity else _launch(x, b, None, None, None, 0, dim, cfg)
        need_x = 'x' in spec.ref or spec.has_2nd_grad
        # 'linear' saves nothing in the reference (ref=''), which makes its CUDA path ignore the clamp in the
        # gradient; the impl='ref' semantics (and the oracle) gate on the clamped output, so keep y then.
        need_y = 'y' in spec.ref or (act == 'linear' and clamp >= 0)
        ctx.save_for_backward(x if need_x else None, b if need_x else None, y if need_y else None)
        ctx.meta = (dim, act, cfg)
        return y

    @staticmethod
    def backward(ctx, dy):
        dim, act, cfg = ctx.meta
        x, b, y = ctx.saved_tensors
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            _, _, gain, clamp = cfg
            dx = dy
            if act != 'linear' or gain != 1 or clamp >= 0:
                dx = _BiasActGrad.apply(dy, x, b, y, dim, act, cfg)
        if ctx.needs_input_grad[1]:
            db = dx.sum([i for i in range(dx.ndim) if i != dim])
        return dx, db, None, None, None


class _BiasActGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, x, b, y, dim, act, cfg):
        spec = activation_funcs[act]
        ref = y if y is not None else x
        dy = _dense_like(dy, ref) if ref is not None else _dense(dy)
        dx = _launch(dy, b, x, y, None, 1, dim, cfg)
        ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
        ctx.meta = (dim, act, cfg)
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        dim, act, cfg = ctx.meta
        spec = activation_funcs[act]
        dy, x, b, y = ctx.saved_tensors
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGrad.apply(d_dx, x, b, y, dim, act, cfg)
        if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = _launch(_dense_like(d_dx, dy), b, x, y, dy, 2, dim, cfg)
        if spec.has_2nd_grad and ctx.needs_input_grad[2]:
            d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
        return d_dy, d_x, d_b, None, None, None, None


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain) for x of any shape (float16/32/64, CUDA).

    `b` is a 1-D tensor of x's dtype matching dimension `dim`; `act` one of `activation_funcs`;
    `alpha` / `gain` default per activation; `clamp=None` disables clamping.  Differentiable to
    second order like the reference (bias_act.py:53-88).  `impl` is kept for signature
    compatibility; both values run the sm_100a kernel.
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        if b.dtype != x.dtype:
            raise TypeError('bias_act: b must have the same dtype as x')
    cfg = (spec.cuda_idx,
           float(alpha if alpha is not None else spec.def_alpha),
           float(gain if gain is not None else spec.def_gain),
           float(clamp if clamp is not None else -1))
    return _BiasAct.apply(x, b, dim, act, cfg)
