"""upfirdn2d -- pad, upsample, FIR-filter and downsample a batch of 2-D images on B200.

Host-side mirror of the reference op `torch_utils/ops/upfirdn2d.py` (public names, argument
meaning, defaults and autograd behaviour are the same: `setup_filter`, `upfirdn2d`, `filter2d`,
`upsample2d`, `downsample2d`), backed by `sg3_upfirdn2d` of libsg3_b200.so instead of the
JIT-built `upfirdn2d_plugin` (upfirdn2d.cpp:16-98).  CUDA tensors only; no CPU path.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import capi

# ---------------------------------------------------------------------------------------------
# Host copies of FIR taps.  The kernels take taps by value (launch parameters), so each filter
# tensor is copied to the host once and remembered for as long as the tensor object lives and is
# not modified in place.

_taps = {}


def _version_of(f):
    """Version counter of `f`, or None for tensors that do not track one (created under torch.inference_mode())."""
    try:
        return f._version
    except RuntimeError:
        return None


def host_taps(f):
    """float32 numpy copy of filter tensor `f`, cached per tensor object + version counter.

    Update a filter with `f.copy_(...)` (bumps the version); writes through `f.data` are invisible to the cache.
    Inference tensors have no version counter and are immutable outside inference mode: they are keyed on
    (data_ptr, shape) instead."""
    key = id(f)
    ver = _version_of(f)
    if ver is None:
        ver = ('inference', f.data_ptr(), tuple(f.shape))
    hit = _taps.get(key)
    if hit is not None and hit[0]() is f and hit[1] == ver:
        return hit[2]
    arr = np.ascontiguousarray(f.detach().to(device='cpu', dtype=torch.float32).numpy())
    _taps[key] = (weakref.ref(f, lambda _r, k=key: _taps.pop(k, None)), ver, arr)
    return arr


def _scaling(v):
    if isinstance(v, int):
        return v, v
    assert isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(q, int) for q in v)
    sx, sy = v
    assert sx >= 1 and sy >= 1
    return sx, sy


def _padding(v):
    if isinstance(v, (int, np.integer)):
        v = [v, v]
    assert isinstance(v, (list, tuple)) and all(isinstance(q, (int, np.integer)) for q in v)
    v = [int(q) for q in v]
    if len(v) == 2:
        v = [v[0], v[0], v[1], v[1]]
    x0, x1, y0, y1 = v
    return x0, x1, y0, y1


def _filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])      # width, height


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Prepare a FIR filter for `upfirdn2d()`; same contract as upfirdn2d.py:71-115.

    Accepts a tensor / array / list of shape [H, W], [taps], [] or None (identity); returns a
    float32 tensor [H, W] (dense) or [taps] (separable; chosen automatically for >= 8 taps).
    """
    f = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    assert f.ndim <= 2 and f.numel() > 0
    if f.ndim == 0:
        f = f.reshape(1)
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


# ---------------------------------------------------------------------------------------------

def upfirdn2d_raw(x, taps2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """One kernel launch: dense 2-D numpy filter `taps2d` [fH, fW]; x any strides; y contiguous."""
    capi.require_cuda(x, 'upfirdn2d')
    assert x.ndim == 4
    fh, fw = taps2d.shape
    n, c, ih, iw = x.shape
    ow = (iw * upx + padx0 + padx1 - fw + downx) // downx
    oh = (ih * upy + pady0 + pady1 - fh + downy) // downy
    if ow < 1 or oh < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')
    if x.numel() == 0:
        raise RuntimeError('upfirdn2d: x has zero size')
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device)
    xs = capi.c_i64x4(*x.stride())
    ys = capi.c_i64x4(*y.stride())
    with torch.cuda.device(x.device):
        rc = capi.lib().sg3_upfirdn2d(
            x.data_ptr(), y.data_ptr(), taps2d.ctypes.data, n, c, ih, iw, oh, ow,
            ctypes.byref(xs), ctypes.byref(ys), fw, fh, upx, upy, downx, downy, padx0, pady0,
            int(bool(flip)), float(gain), capi.dtype_code(x.dtype), capi.stream_ptr(x.device))
    capi.check(rc, 'sg3_upfirdn2d')
    return y


def upfirdn2d_sep_raw(x, taps_x, taps_y, up, down, padx0, padx1, pady0, pady1, flip, gain):
    """Separable filter (1-D numpy taps per axis) in one launch (sg3_upfirdn2d_sep): the warp-streaming kernel for fp32 with
    up 2 / down 2 / neither, the tiled one-pass kernel otherwise.  Returns None when the library has no kernel for the shape."""
    capi.require_cuda(x, 'upfirdn2d')
    fw, fh = int(taps_x.shape[0]), int(taps_y.shape[0])
    n, c, ih, iw = x.shape
    ow = (iw * up + padx0 + padx1 - fw + down) // down
    oh = (ih * up + pady0 + pady1 - fh + down) // down
    if ow < 1 or oh < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')
    if x.numel() == 0 or x.dtype not in (torch.float16, torch.float32) or x.stride(3) != 1:
        return None
    y = torch.empty([n, c, oh, ow], dtype=x.dtype, device=x.device)
    xs = capi.c_i64x4(*x.stride())
    ys = capi.c_i64x4(*y.stride())
    with torch.cuda.device(x.device):
        rc = capi.lib().sg3_upfirdn2d_sep(
            x.data_ptr(), y.data_ptr(), taps_x.ctypes.data, taps_y.ctypes.data, n, c, ih, iw, oh, ow,
            ctypes.byref(xs), ctypes.byref(ys), fw, fh, up, down, padx0, pady0,
            int(bool(flip)), float(gain), capi.dtype_code(x.dtype), capi.stream_ptr(x.device))
    if rc == capi.SG3_E_NOKERNEL:
        return None
    capi.check(rc, 'sg3_upfirdn2d_sep')
    return y


_rank1 = {}


def _rank1_factors(taps2d):
    """(taps_x, taps_y) when the dense filter is an outer product to fp32 rounding, else None.  `setup_filter` turns short 1-D
    filters (fewer than 8 taps, e.g. the [1, 3, 3, 1] blur) into their dense outer product (upfirdn2d.py:103-105): running those as
    two 1-D passes in one kernel costs 2 * taps instead of taps^2 multiply-adds per output and the same bytes."""
    key = (taps2d.shape, taps2d.tobytes())
    hit = _rank1.get(key)
    if hit is None:
        hit = False
        fh, fw = taps2d.shape
        if fh > 1 and fw > 1:
            a = taps2d.astype(np.float64)
            i, j = np.unravel_index(np.argmax(np.abs(a)), a.shape)
            if a[i, j] != 0:
                ty, tx = a[:, j], a[i, :] / a[i, j]
                if np.max(np.abs(np.outer(ty, tx) - a)) <= 4e-7 * np.abs(a[i, j]):
                    hit = (np.ascontiguousarray(tx, dtype=np.float32), np.ascontiguousarray(ty, dtype=np.float32))
        if len(_rank1) > 256:
            _rank1.clear()
        _rank1[key] = hit
    return hit or None


_ONE = np.ones((1, 1), np.float32)


def _run(x, f, upx, upy, downx, downy, pads, flip, gain):
    """Dense filter -> one launch; separable filter -> the fused one-pass kernel, else x pass then y pass (upfirdn2d.py:241-246)."""
    px0, px1, py0, py1 = pads
    if f is None:
        return upfirdn2d_raw(x, _ONE, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain)
    taps = host_taps(f)
    if taps.ndim == 1 and taps.shape[0] == 1:
        taps = (taps * taps).reshape(1, 1)
    if taps.ndim == 2:
        sep = _rank1_factors(taps) if (upx == upy and downx == downy) else None
        if sep is not None:                    # a dense outer product (setup_filter of a short 1-D filter): two 1-D passes, one launch
            y = upfirdn2d_sep_raw(x, sep[0], sep[1], upx, downx, px0, px1, py0, py1, flip, gain)
            if y is not None:
                return y
        return upfirdn2d_raw(x, taps, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain)
    if upx == upy and downx == downy:          # one pass over HBM instead of the reference's two launches
        y = upfirdn2d_sep_raw(x, taps, taps, upx, downx, px0, px1, py0, py1, flip, gain)
        if y is not None:
            return y
    y = upfirdn2d_raw(x, taps.reshape(1, -1), upx, 1, downx, 1, px0, px1, 0, 0, flip, 1.0)
    return upfirdn2d_raw(y, taps.reshape(-1, 1), 1, upy, 1, downy, 0, 0, py0, py1, flip, gain)


class _Upfirdn2d(torch.autograd.Function):
    """Differentiable (any order) in x; the adjoint is the same op with up/down swapped, the filter
    flipped and the padding of upfirdn2d.py:257-262."""

    @staticmethod
    def forward(ctx, x, f, cfg):
        upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = cfg
        y = _run(x, f, upx, upy, downx, downy, (px0, px1, py0, py1), flip, gain)
        ctx.f = f
        ctx.cfg = cfg
        ctx.x_hw = (x.shape[2], x.shape[3])
        return y

    @staticmethod
    def backward(ctx, dy):
        upx, upy, downx, downy, px0, px1, py0, py1, flip, gain = ctx.cfg
        assert not ctx.needs_input_grad[1], 'upfirdn2d: the filter is not differentiable'
        dx = None
        if ctx.needs_input_grad[0]:
            ih, iw = ctx.x_hw
            oh, ow = dy.shape[2], dy.shape[3]
            fw, fh = _filter_size(ctx.f)
            adj = (downx, downy, upx, upy,
                   fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1,
                   fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1,
                   not flip, gain)
            dx = _Upfirdn2d.apply(dy, ctx.f, adj)
        return dx, None, None


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, upsample, filter and downsample `x` [N, C, H, W] (float16/32/64, CUDA).

    Steps per channel: insert `up-1` zeros after each pixel; pad (negative = crop) by `padding`
    = int | [x, y] | [x0, x1, y0, y1]; convolve with `f` ([H, W], [taps] separable, or None),
    keeping only fully covered outputs; keep every `down`-th pixel.  `flip_filter=False` is a true
    convolution, True a correlation.  `impl` is accepted for signature compatibility: both 'cuda'
    and 'ref' run the sm_100a kernel (the reference's slow PyTorch path lives only in oracle/).
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ('ref', 'cuda')
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32)
    upx, upy = _scaling(up)
    downx, downy = _scaling(down)
    cfg = (upx, upy, downx, downy) + _padding(padding) + (bool(flip_filter), float(gain))
    return _Upfirdn2d.apply(x, f, cfg)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR-filter keeping the spatial size (upfirdn2d.py:278-310)."""
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample by `up`; output size is a multiple of the input (upfirdn2d.py:314-349)."""
    upx, upy = _scaling(up)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Downsample by `down`; output size is a fraction of the input (upfirdn2d.py:353-388)."""
    downx, downy = _scaling(down)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
