"""GPU parity of the op layer (through the C ABI) against the golden vectors generated from the
reference and against the CPU oracle on seeded inputs.

Tolerance: BASELINE.json's north_star allows 1e-3 max relative error in fp32; these single-op
checks hold 2e-5 (max |err| / max |ref|).  fp16 I/O rounds once at the store: 2e-3.
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

TOL32 = 2e-5
TOL16 = 2e-3


@pytest.fixture(scope='module')
def ops():
    import sg3_b200
    assert torch.cuda.is_available()
    return sg3_b200


def cu(a, grad=False, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if dtype is not None:
        t = t.to(dtype)
    return t.requires_grad_(grad)


def _opt(c, k):
    return c[k] if k in c else None


def _f(c, k):
    return None if k not in c else cu(c[k])


@pytest.mark.parametrize('case', golden('ops.npz').cases('upfirdn'))
def test_upfirdn2d_golden(ops, case):
    c = golden('ops.npz').case('upfirdn', case)
    x = cu(c['x'], True)
    y = ops.upfirdn2d.upfirdn2d(x, _f(c, 'f'), up=[int(v) for v in c['up']], down=[int(v) for v in c['down']],
                                padding=[int(v) for v in c['padding']], flip_filter=bool(c['flip']), gain=float(c['gain']))
    assert tuple(y.shape) == c['y'].shape
    assert rel_err(y.detach().cpu().numpy(), c['y']) < TOL32
    (dx,) = torch.autograd.grad(y, x, cu(c['dy']))
    assert rel_err(dx.cpu().numpy(), c['dx']) < TOL32


@pytest.mark.parametrize('case', golden('ops.npz').cases('bias_act'))
def test_bias_act_golden(ops, case):
    c = golden('ops.npz').case('bias_act', case)
    act = case.rsplit('_', 1)[0]
    kw = dict(dim=int(c['dim']), act=act,
              alpha=None if np.isnan(c['alpha']) else float(c['alpha']),
              gain=None if np.isnan(c['gain']) else float(c['gain']),
              clamp=None if c['clamp'] < 0 else float(c['clamp']))
    x, b, dy = cu(c['x'], True), cu(c['b'], True), cu(c['dy'], True)
    y = ops.bias_act.bias_act(x, b, **kw)
    assert rel_err(y.detach().cpu().numpy(), c['y']) < TOL32
    dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
    assert rel_err(dx.detach().cpu().numpy(), c['dx']) < TOL32
    assert rel_err(db.detach().cpu().numpy(), c['db']) < TOL32
    g2 = torch.autograd.grad(dx, [dy, x], cu(c['ddx']), allow_unused=True)
    assert rel_err(g2[0].cpu().numpy(), c['d_dy']) < TOL32
    if np.abs(c['d_x']).max() > 0:
        assert rel_err(g2[1].cpu().numpy(), c['d_x']) < 1e-4
    else:
        assert g2[1] is None or float(g2[1].abs().max()) == 0


@pytest.mark.parametrize('impl', ['cuda', 'ref'])
@pytest.mark.parametrize('case', golden('ops.npz').cases('flrelu'))
def test_filtered_lrelu_golden(ops, case, impl):
    c = golden('ops.npz').case('flrelu', case)
    kw = dict(up=int(c['up']), down=int(c['down']), padding=[int(v) for v in c['padding']], gain=float(c['gain']),
              slope=float(c['slope']), clamp=None if c['clamp'] < 0 else float(c['clamp']), flip_filter=bool(c['flip']))
    x, b = cu(c['x'], True), cu(c['b'], True)
    ops.filtered_lrelu._quiet_fallback = True
    y = ops.filtered_lrelu.filtered_lrelu(x, _f(c, 'fu'), _f(c, 'fd'), b, impl=impl, **kw)
    assert tuple(y.shape) == c['y'].shape
    assert rel_err(y.detach().cpu().numpy(), c['y']) < TOL32
    dx, db = torch.autograd.grad(y, [x, b], cu(c['dy']))
    assert rel_err(dx.cpu().numpy(), c['dx']) < 5e-5
    assert rel_err(db.cpu().numpy(), c['db']) < 5e-5


@pytest.mark.parametrize('case', ['R_same', 'R_up4', 'T_same', 'T_up4', 'crit_last', 'torgb'])
def test_filtered_lrelu_fp16(ops, case):
    """fp16 I/O (the reference's default for L5-L14): fp32 accumulate, one rounding at the store."""
    from oracle import sg3_oracle as orc
    c = golden('ops.npz').case('flrelu', case)
    kw = dict(up=int(c['up']), down=int(c['down']), padding=[int(v) for v in c['padding']], gain=float(c['gain']),
              slope=float(c['slope']), clamp=None if c['clamp'] < 0 else float(c['clamp']), flip_filter=bool(c['flip']))
    x16 = c['x'].astype(np.float16)
    b16 = c['b'].astype(np.float16)
    ref = orc.filtered_lrelu(x16.astype(np.float32), _opt(c, 'fu'), _opt(c, 'fd'), b16.astype(np.float32), **kw)
    ops.filtered_lrelu._quiet_fallback = True
    y = ops.filtered_lrelu.filtered_lrelu(cu(x16), _f(c, 'fu'), _f(c, 'fd'), cu(b16), **kw)
    assert y.dtype == torch.float16
    assert rel_err(y.float().cpu().numpy(), ref) < TOL16


def test_filtered_lrelu_double_backward(ops):
    """Gradients of any order: d/dx of <dx, v> must match the oracle's linear adjoint applied twice."""
    from oracle import sg3_oracle as orc
    c = golden('ops.npz').case('flrelu', 'T_same')
    kw = dict(up=2, down=2, padding=[int(v) for v in c['padding']], gain=float(c['gain']), slope=float(c['slope']),
              clamp=float(c['clamp']), flip_filter=False)
    x, dy = cu(c['x'], True), cu(c['dy'], True)
    ops.filtered_lrelu._quiet_fallback = True
    y = ops.filtered_lrelu.filtered_lrelu(x, _f(c, 'fu'), _f(c, 'fd'), None, **kw)
    (dx,) = torch.autograd.grad(y, x, dy, create_graph=True)
    v = torch.randn_like(dx)
    (d_dy,) = torch.autograd.grad(dx, dy, v)
    # dx is linear in dy given the signs, so d<dx,v>/d(dy) = forward-linearised op applied to v:
    _, signs = orc.filtered_lrelu(c['x'], c['fu'], c['fd'], None, return_signs=True, **kw)
    a = orc.upfirdn2d(v.cpu().numpy(), c['fu'], up=2, padding=kw['padding'], gain=4)
    a = np.ascontiguousarray(a)
    orc.lrelu_act_(a, kw['gain'], kw['slope'], None, signs, 0, 0, mode=2)
    ref = orc.upfirdn2d(a, c['fd'], down=2)
    assert rel_err(d_dy.cpu().numpy(), ref) < 5e-5


def test_strided_and_large_vs_oracle(ops):
    """Non-contiguous input views, odd sizes, many planes."""
    from oracle import sg3_oracle as orc
    g = golden('ops.npz').case('flrelu', 'R_same')
    rng = np.random.RandomState(7)
    base = rng.randn(3, 10, 45, 53).astype(np.float32) * 3
    xb = cu(base)
    xv = xb[:, 1:9:2, 2:43, 5:50]                 # strided view, [3,4,41,45]
    b = rng.randn(4).astype(np.float32)
    kw = dict(up=2, down=2, padding=[11, 10, 11, 10], gain=np.sqrt(2), slope=0.2, clamp=4.0)
    ops.filtered_lrelu._quiet_fallback = True
    y = ops.filtered_lrelu.filtered_lrelu(xv, cu(g['fu']), cu(g['fd']), cu(b), **kw)
    ref = orc.filtered_lrelu(np.ascontiguousarray(base[:, 1:9:2, 2:43, 5:50]), g['fu'], g['fd'], b, **kw)
    assert rel_err(y.cpu().numpy(), ref) < TOL32


def test_errors(ops):
    x = torch.zeros(1, 2, 4, 4, device='cuda')
    f = torch.ones(12, device='cuda')
    with pytest.raises(RuntimeError):
        ops.filtered_lrelu.filtered_lrelu(x, f, f, None, up=2, down=2, padding=0)      # up buffer < down filter
    with pytest.raises(RuntimeError):
        ops.filtered_lrelu.filtered_lrelu(torch.zeros(1, 2, 4, 4), None, None, None)     # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ops.bias_act.bias_act(x, torch.zeros(2, device='cuda', dtype=torch.float16))


# ---------------------------------------------------------------------------------------------
# The fused warp-streaming kernel (flrelu_stream.cuh): multi-strip / multi-chunk shapes, every
# sign mode, both dtypes -- against the oracle, and proof that the fused path (not the generic
# composition) produced the result.

def _design(up_taps, radial_down):
    from oracle import sg3_oracle as orc
    fu = orc.design_lowpass_filter(up_taps, 11.3, 26.0, 128 if up_taps == 12 else 256)
    fd = orc.design_lowpass_filter(12, 16.0, 36.0, 128, radial=radial_down)
    return fu, fd


FUSED_CASES = [
    # name, up, radial, N, C, in, padding
    ('R_same_148', 2, True, 1, 3, 148, [11, 10, 11, 10]),       # 3 strips wide, row chunks
    ('R_up4_84', 4, True, 1, 2, 84, [-2, -5, -2, -5]),
    ('T_same_150', 2, False, 2, 2, 150, [9, 8, 9, 8]),
    ('T_up4_86', 4, False, 1, 3, 86, [-6, -9, -6, -9]),
    ('crit_70', 2, False, 1, 2, 70, [-9, -10, -9, -10]),
    ('odd_pad_phase', 4, True, 1, 2, 37, [3, 2, 1, 4]),         # exercises ex, ey = 1..3
    ('odd_pad_phase2', 2, False, 1, 2, 41, [6, 7, 9, 4]),
    ('many_planes', 2, True, 4, 40, 36, [11, 10, 11, 10]),
    ('full_last_strip', 2, False, 1, 2, 84, [-9, -10, -9, -10]),   # out 64: the last strip is exactly 64 wide
    ('two_full_strips', 2, True, 1, 2, 128, [11, 10, 11, 10]),      # out 128
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('case', FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_stream_forward(ops, case, dtype):
    from oracle import sg3_oracle as orc
    name, up, radial, N, C, size, pad = case
    fu, fd = _design(6 * up, radial)
    rng = np.random.RandomState(hash(name) % 1000)
    x = (rng.randn(N, C, size, size + 3) * 3).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    if dtype == torch.float16:
        x, b = x.astype(np.float16).astype(np.float32), b.astype(np.float16).astype(np.float32)
    kw = dict(up=up, down=2, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=6.0)
    ref, ref_signs = orc.filtered_lrelu(x, fu, fd, b, return_signs=True, **kw)
    fl = ops.filtered_lrelu
    cfg = (up, 2) + tuple(pad) + (float(np.sqrt(2)), 0.2, 6.0, False)
    xt, bt = cu(x, dtype=dtype), cu(b, dtype=dtype)
    for write in (False, True):
        res = fl._fused(xt, cu(fu), cu(fd), bt, None, 0, 0, cfg, write)
        assert res is not None, 'fused kernel must support this configuration'
        y, so = res
        assert tuple(y.shape) == ref.shape
        assert rel_err(y.float().cpu().numpy(), ref) < (TOL32 if dtype == torch.float32 else TOL16)
        if write:
            assert tuple(so.shape) == ref_signs.shape
            sw_active = 2 * ref.shape[3] + 10
            got = so.cpu().numpy()
            # unpack 2-bit codes over the active width and compare; a code may differ only where the
            # pre-activation sits within rounding of 0 or of the clamp
            def unpack(s):
                s = s.astype(np.uint8)
                return np.stack([(s >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(*s.shape[:3], -1)[..., :sw_active]
            diff = unpack(got) != unpack(ref_signs)
            assert diff.sum() <= (2 if dtype == torch.float32 else 5e-3 * diff.size), (int(diff.sum()), np.argwhere(diff)[:8])


@pytest.mark.parametrize('case', ['T_same_150', 'crit_70', 'odd_pad_phase2'])
def test_fused_stream_backward_signs(ops, case):
    """up2/down2 separable layers are self-adjoint in shape: backward runs the fused kernel in sign-READ mode."""
    from oracle import sg3_oracle as orc
    name, up, radial, N, C, size, pad = [c for c in FUSED_CASES if c[0] == case][0]
    fu, fd = _design(12, False)
    rng = np.random.RandomState(11)
    x = (rng.randn(N, C, size, size) * 3).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    kw = dict(up=2, down=2, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=6.0)
    y_ref, signs = orc.filtered_lrelu(x, fu, fd, b, return_signs=True, **kw)
    dy = rng.randn(*y_ref.shape).astype(np.float32)
    kwb = dict(kw)
    kwb.pop('clamp')
    dx_ref, db_ref = orc.filtered_lrelu_bwd(dy, signs, x.shape, fu, fd, **kwb)
    xt, bt = cu(x, True), cu(b, True)
    calls = []
    orig = ops.filtered_lrelu._fused

    def spy(*a, **k):
        r = orig(*a, **k)
        calls.append(r is not None)
        return r
    ops.filtered_lrelu._fused = spy
    try:
        y = ops.filtered_lrelu.filtered_lrelu(xt, cu(fu), cu(fd), bt, **kw)
        dx, db = torch.autograd.grad(y, [xt, bt], cu(dy))
    finally:
        ops.filtered_lrelu._fused = orig
    assert calls == [True, True], calls            # forward (sign write) and backward (sign read) both fused
    assert rel_err(dx.cpu().numpy(), dx_ref) < 5e-5
    assert rel_err(db.cpu().numpy(), db_ref) < 5e-5


def test_fused_stream_full_size_properties(ops):
    """Size-independent properties at a BASELINE-size layer (R-1024 L11: 102ch 1044^2): linearity of the
    op in the unclamped positive regime and exact zero-response to zero input with zero bias."""
    from oracle import sg3_oracle as orc
    fu, fd = _design(12, True)
    fl = ops.filtered_lrelu
    x = torch.rand(1, 6, 1044, 1044, device='cuda') + 0.5          # strictly positive -> lrelu is identity
    kw = dict(up=2, down=2, padding=[11, 10, 11, 10], gain=1.0, slope=0.2, clamp=None)
    y1 = fl.filtered_lrelu(x, cu(fu), cu(fd), None, **kw)
    y2 = fl.filtered_lrelu(2 * x, cu(fu), cu(fd), None, **kw)
    assert tuple(y1.shape) == (1, 6, 1044, 1044)
    inner = (slice(None), slice(None), slice(16, -16), slice(16, -16))   # borders see the zero padding ripple (may go negative)
    assert float((y2[inner] - 2 * y1[inner]).abs().max()) < 1e-4 * float(y1.abs().max())
    # DC gain of both low-pass filters is 1: the interior of a constant image stays constant
    yc = fl.filtered_lrelu(torch.ones(1, 2, 1044, 1044, device='cuda'), cu(fu), cu(fd), None, **kw)
    assert float((yc[inner] - 1).abs().max()) < 1e-4
    z = fl.filtered_lrelu(torch.zeros(1, 2, 1044, 1044, device='cuda'), cu(fu), cu(fd), None, **kw)
    assert float(z.abs().max()) == 0.0
    # one plane against the oracle at full size
    ref = orc.filtered_lrelu(x[:, :1].cpu().numpy(), fu, fd, None, **kw)
    assert rel_err(y1[:, :1].cpu().numpy(), ref) < TOL32


BWD_CASES = [
    # name, up (forward), N, C, in, padding: forward = separable up -> dense radial down 2; backward = dense up 2 -> separable down `up`
    ('R_same_148', 2, 1, 3, 148, [11, 10, 11, 10]),
    ('R_up4_84', 4, 1, 2, 84, [-2, -5, -2, -5]),
    ('R_same_odd', 2, 2, 2, 45, [10, 11, 9, 12]),
    ('R_up4_odd', 4, 1, 3, 37, [3, 2, 1, 4]),
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('case', BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_fused_dense_up_backward(ops, case, dtype):
    """Backward of the config-R layers runs flrelu_bwd_stream (dense up 2 -> separable down 2 / 4, sign READ):
    dx, db against the oracle's adjoint, and proof that both directions took a fused kernel."""
    from oracle import sg3_oracle as orc
    name, up, N, C, size, pad = case
    fu, fd = _design(6 * up, True)
    rng = np.random.RandomState(len(name) * 7)
    x = (rng.randn(N, C, size, size + 4) * 3).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    if dtype == torch.float16:
        x, b = x.astype(np.float16).astype(np.float32), b.astype(np.float16).astype(np.float32)
    kw = dict(up=up, down=2, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=6.0)
    y_ref, signs = orc.filtered_lrelu(x, fu, fd, b, return_signs=True, **kw)
    dy = rng.randn(*y_ref.shape).astype(np.float32)
    if dtype == torch.float16:
        dy = dy.astype(np.float16).astype(np.float32)
    kwb = dict(kw)
    kwb.pop('clamp')
    dx_ref, db_ref = orc.filtered_lrelu_bwd(dy, signs, x.shape, fu, fd, **kwb)
    xt, bt = cu(x, True, dtype=dtype), cu(b, True, dtype=dtype)
    calls = []
    orig = ops.filtered_lrelu._fused

    def spy(*a, **k):
        r = orig(*a, **k)
        calls.append(r is not None)
        return r
    ops.filtered_lrelu._fused = spy
    try:
        y = ops.filtered_lrelu.filtered_lrelu(xt, cu(fu), cu(fd), bt, **kw)
        dx, db = torch.autograd.grad(y, [xt, bt], cu(dy, dtype=dtype))
    finally:
        ops.filtered_lrelu._fused = orig
    assert calls == [True, True], calls
    tol = 5e-5 if dtype == torch.float32 else 4e-3     # fp16: the forward's fp16 rounding moves a few sign decisions
    assert rel_err(dx.float().cpu().numpy(), dx_ref) < tol
    assert rel_err(db.float().cpu().numpy(), db_ref) < (5e-5 if dtype == torch.float32 else 2e-2)


def test_fused_dense_up_forward_and_sep_down4(ops):
    """The same kernel family as a forward op: dense up 2 -> separable down 2, and separable up 2 -> down 4."""
    from oracle import sg3_oracle as orc
    fu12, fdR = _design(12, True)
    fu24, _ = _design(24, True)
    rng = np.random.RandomState(3)
    x = (rng.randn(2, 3, 70, 66) * 2).astype(np.float32)
    b = rng.randn(3).astype(np.float32)
    fl = ops.filtered_lrelu
    for fu, fd, down, pad in [(fdR, fu12, 2, [10, 11, 10, 11]), (fu12, fu24, 4, [7, 9, 12, 6]), (fdR, fu24, 4, [16, 15, 17, 14])]:
        kw = dict(up=2, down=down, padding=pad, gain=1.3, slope=0.1, clamp=2.0, flip_filter=True)
        ref = orc.filtered_lrelu(x, fu, fd, b, **kw)
        cfg = (2, down) + tuple(pad) + (1.3, 0.1, 2.0, True)
        res = fl._fused(cu(x), cu(fu), cu(fd), cu(b), None, 0, 0, cfg, False)
        assert res is not None
        assert rel_err(res[0].cpu().numpy(), ref) < TOL32


# ---------------------------------------------------------------------------------------------
# upfirdn2d fast paths (1-D polyphase kernels, small dense filters) vs the CPU oracle.

UPFIRDN_FAST = [
    # name, taps (1-D -> separable two-pass; 2-D -> dense), up, down, padding, flip, shape
    ('sep12_up2', 12, 2, 1, [11, 10, 11, 10], False, (2, 3, 37, 45)),
    ('sep24_up4', 24, 4, 1, [-2, -5, -2, -5], False, (1, 4, 29, 33)),
    ('sep12_down2', 12, 1, 2, 0, False, (2, 3, 70, 86)),
    ('sep24_down4', 24, 1, 4, [3, 1, 2, 0], True, (1, 2, 90, 101)),
    ('sep8_same', 8, 1, 1, [4, 3, 4, 3], True, (2, 2, 33, 65)),
    ('sep12_up2_crop', 12, 2, 1, [-3, 7, 5, -4], False, (1, 2, 40, 40)),
    ('dense4_filter', (4, 4), 1, 1, [2, 1, 2, 1], False, (2, 3, 50, 67)),
    ('dense4_up2', (4, 4), 2, 1, [2, 1, 2, 1], False, (2, 3, 31, 35)),
    ('dense4_down2', (4, 4), 1, 2, [1, 1, 1, 1], True, (2, 3, 64, 70)),
    ('dense3x4_up4', (3, 4), 4, 1, [5, 2, 3, 1], False, (1, 2, 17, 19)),
    ('dense8_up2', (8, 8), 2, 1, [4, 3, 4, 3], False, (1, 2, 21, 23)),
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('case', UPFIRDN_FAST, ids=[c[0] for c in UPFIRDN_FAST])
def test_upfirdn2d_fast_paths(ops, case, dtype):
    from oracle import sg3_oracle as orc
    name, taps, up, down, padding, flip, shape = case
    rng = np.random.RandomState(len(name))
    f = rng.randn(*((taps,) if isinstance(taps, int) else taps)).astype(np.float32)
    x = rng.randn(*shape).astype(np.float32)
    if dtype == torch.float16:
        x = x.astype(np.float16).astype(np.float32)
    before = ops.capi.lib().sg3_launch_count()
    y = ops.upfirdn2d.upfirdn2d(cu(x).to(dtype), cu(f), up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    assert ops.capi.lib().sg3_launch_count() - before == 1, 'separable filters run the one-pass kernel, dense ones a single launch'
    ref = orc.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    assert y.shape == ref.shape
    assert rel_err(y.float().cpu().numpy(), ref) < (TOL32 if dtype == torch.float32 else 2e-3)
    if isinstance(taps, int):       # the two-launch 1-D kernels (what other up/down combinations use) agree as well
        px0, px1, py0, py1 = ops.upfirdn2d._padding(padding)
        t = ops.upfirdn2d.host_taps(cu(f))
        y1 = ops.upfirdn2d.upfirdn2d_raw(cu(x).to(dtype), t.reshape(1, -1), up, 1, down, 1, px0, px1, 0, 0, flip, 1.0)
        y2 = ops.upfirdn2d.upfirdn2d_raw(y1, t.reshape(-1, 1), 1, up, 1, down, 0, 0, py0, py1, flip, up * up)
        assert rel_err(y2.float().cpu().numpy(), ref) < (TOL32 if dtype == torch.float32 else 4e-3)


# ---------------------------------------------------------------------------------------------
# The warp-streaming separable kernel (csrc/upfirdn2d_stream.cu): every alignment case is folded into its tap tables, so sweep
# the padding residues (both parities / all four residues mod 4 per axis), row alignments (16- and 8-byte rows), tap counts of every
# instantiation, crops, more than one strip (128 output columns) and more than one row chunk, views with a plane / row pitch.

UPFIRDN_STREAM = [
    # name, taps, up, down, padding, flip, shape, (row pitch, column offset) of the input view
    ('up2_t12', 12, 2, 1, [11, 10, 11, 10], False, (2, 3, 36, 44), None),
    ('up2_t12_odd_pad', 12, 2, 1, [10, 11, 12, 9], True, (1, 2, 40, 152), None),        # 304 output columns: 3 strips
    ('up2_t4', 4, 2, 1, [2, 1, 2, 1], False, (2, 2, 70, 36), None),
    ('up2_t8_crop', 8, 2, 1, [-3, 7, 5, -4], False, (1, 2, 300, 40), None),             # 5 row chunks of 128
    ('up2_t24', 24, 2, 1, [23, 22, 21, 20], False, (1, 2, 34, 38), None),
    ('up2_t11_8byte_rows', 11, 2, 1, [5, 5, 6, 4], False, (1, 3, 33, 42), None),         # rows 8-byte aligned only
    ('down2_t12', 12, 1, 2, 0, False, (2, 3, 70, 88), None),
    ('down2_t12_pads', 12, 1, 2, [1, 2, 3, 0], True, (1, 2, 90, 600), None),             # 296 output columns
    ('down2_t12_pad2', 12, 1, 2, [2, 3, 1, 5], False, (1, 2, 64, 72), None),
    ('down2_t12_pad3', 12, 1, 2, [3, 0, 2, 2], False, (1, 2, 64, 72), None),
    ('down2_t4', 4, 1, 2, [1, 1, 1, 1], True, (2, 3, 64, 72), None),
    ('down2_t8_rows', 8, 1, 2, [4, 3, 4, 3], False, (1, 2, 700, 36), None),              # 3 row chunks
    ('down2_t12_2098', 12, 1, 2, 0, False, (1, 1, 150, 2098), None),                     # the reference-path shape: 8-byte rows
    ('same_t12', 12, 1, 1, [6, 5, 6, 5], False, (2, 2, 33, 64), None),
    ('same_t8_flip', 8, 1, 1, [4, 3, 4, 3], True, (2, 2, 33, 68), None),
    ('same_t4_crop', 4, 1, 1, [-2, 4, 3, -1], False, (1, 2, 40, 260), None),
    ('same_t5', 5, 1, 1, [2, 2, 2, 2], False, (1, 2, 40, 48), None),
    ('up2_t12_pitched_view', 12, 2, 1, [11, 10, 11, 10], False, (2, 3, 36, 44), (52, 4)),
    ('down2_t12_pitched_view', 12, 1, 2, [0, 0, 0, 0], False, (2, 2, 70, 86), (96, 2)),   # 8-byte aligned view
]


@pytest.mark.parametrize('case', UPFIRDN_STREAM, ids=[c[0] for c in UPFIRDN_STREAM])
def test_upfirdn2d_stream_kernel(ops, case):
    from oracle import sg3_oracle as orc
    name, taps, up, down, padding, flip, shape, view = case
    rng = np.random.RandomState(len(name) + taps)
    f = rng.randn(taps).astype(np.float32)
    x = rng.randn(*shape).astype(np.float32)
    xt = cu(x)
    if view is not None:                       # [N, C, H, W] window of a wider buffer: row pitch != W, base off the buffer start
        pitch, ofs = view
        buf = torch.full([shape[0], shape[1], shape[2], pitch], float('nan'), device='cuda')
        buf[..., ofs:ofs + shape[3]] = xt
        xt = buf[..., ofs:ofs + shape[3]]
    before = ops.capi.lib().sg3_launch_count()
    y = ops.upfirdn2d.upfirdn2d(xt, cu(f), up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    assert ops.capi.lib().sg3_launch_count() - before == 1
    ref = orc.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    assert y.shape == ref.shape
    assert rel_err(y.cpu().numpy(), ref) < TOL32


@pytest.mark.parametrize('seed', list(range(24)))
def test_upfirdn2d_stream_random_sweep(ops, seed):
    """Random separable cases through the public op: factor (up 2 / down 2 / same rate), 1-24 taps, paddings from crops to wide
    borders, odd heights, widths of every alignment class (16-byte rows -> TMA or cp.async 16, 8-byte rows -> cp.async 8, odd rows ->
    the tiled kernel), one or several strips / row chunks, views with a row pitch -- all against the oracle."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(1000 + seed)
    mode = seed % 3
    up, down = ((2, 1), (1, 2), (1, 1))[mode]
    taps = int(rng.randint(1, 25 if up == 2 else 13))
    n, c = int(rng.randint(1, 3)), int(rng.randint(1, 4))
    h = int(rng.randint(20, 140)) if seed % 4 else int(rng.randint(260, 420))
    w = int(rng.choice([rng.randint(5, 40) * 4, rng.randint(10, 80) * 2, rng.randint(21, 160)]))
    if seed % 5 == 0:
        w = int(rng.randint(70, 110)) * 4                      # several strips
    lo = -min(6, (min(h, w) * up) // 4)
    pad = [int(v) for v in rng.randint(lo, 16, size=4)]
    while (w * up + pad[0] + pad[1] - taps) < down or (h * up + pad[2] + pad[3] - taps) < down:
        pad = [v + 2 for v in pad]
    f = rng.randn(taps).astype(np.float32)
    x = rng.randn(n, c, h, w).astype(np.float32)
    xt = cu(x)
    if seed % 6 == 1:                                          # a view with a row pitch and an offset base
        pitch, ofs = w + int(rng.randint(1, 9)), int(rng.randint(0, 3))
        buf = torch.full([n, c, h, pitch + 4], float('nan'), device='cuda')
        buf[..., ofs:ofs + w] = xt
        xt = buf[..., ofs:ofs + w]
    flip = bool(seed & 1)
    gain = float(rng.choice([1.0, up * up, 0.37]))
    y = ops.upfirdn2d.upfirdn2d(xt, cu(f), up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    ref = orc.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    assert y.shape == ref.shape, (up, down, taps, pad, x.shape)
    assert rel_err(y.cpu().numpy(), ref) < TOL32, (up, down, taps, pad, x.shape)


def test_upfirdn2d_stream_kernel_cp_async_staging():
    """Tensors TMA can address stage their rows by TMA; SG3_UPFIRDN_NO_TMA=1 (read once per process) keeps the 16-byte cp.async
    staging for them: same results.  Runs the aligned cases of the table above in a child process."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import sg3_b200
from oracle import sg3_oracle as orc
sys.path.insert(0, %r)
from test_ops_gpu import UPFIRDN_STREAM, cu
from conftest import rel_err
n = 0
for name, taps, up, down, padding, flip, shape, view in UPFIRDN_STREAM:
    if view is not None or shape[3] %% 4 != 0:
        continue
    rng = np.random.RandomState(len(name) + taps)
    f = rng.randn(taps).astype(np.float32)
    x = rng.randn(*shape).astype(np.float32)
    y = sg3_b200.upfirdn2d.upfirdn2d(cu(x), cu(f), up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    ref = orc.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip, gain=up * up)
    assert rel_err(y.cpu().numpy(), ref) < 2e-5, name
    n += 1
print('cases', n)
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SG3_UPFIRDN_NO_TMA='1')
    res = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    assert 'cases' in res.stdout and int(res.stdout.strip().split()[-1]) >= 8


def test_upfirdn2d_rank1_dense_filters_run_separable(ops):
    """`setup_filter([1, 3, 3, 1])` returns the dense 4 x 4 outer product (upfirdn2d.py:103-105); the op factors it back and runs
    the separable kernel: filter2d / upsample2d / downsample2d with their default paddings, against the oracle's dense evaluation."""
    from oracle import sg3_oracle as orc
    f = ops.upfirdn2d.setup_filter([1, 3, 3, 1], device='cuda')
    assert f.ndim == 2
    assert ops.upfirdn2d._rank1_factors(ops.upfirdn2d.host_taps(f)) is not None
    rng = np.random.RandomState(5)
    x = rng.randn(2, 3, 48, 56).astype(np.float32)
    fh = f.cpu().numpy()
    for fn, kw, okw in ((ops.upfirdn2d.filter2d, {}, dict(up=1, down=1, padding=[2, 1, 2, 1])),
                        (ops.upfirdn2d.upsample2d, dict(up=2), dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4)),
                        (ops.upfirdn2d.downsample2d, dict(down=2), dict(up=1, down=2, padding=[1, 1, 1, 1]))):
        y = fn(cu(x), f, **kw)
        ref = orc.upfirdn2d(x, fh, **okw)
        assert y.shape == ref.shape
        assert rel_err(y.cpu().numpy(), ref) < TOL32
    g = rng.randn(4, 4).astype(np.float32)                     # a full-rank filter stays dense
    assert ops.upfirdn2d._rank1_factors(g) is None


def test_fused_backward_accumulates_bias_gradient(ops):
    """desc.ysum: the sign-READ kernels add the per-channel sum of their output (db of the backward pass) with fp32 atomics;
    the autograd path then skips the reduction over dx.  Compared with dx.sum of the create_graph path (torch reduction)."""
    for up, radial in ((2, True), (4, True), (2, False)):
        fu, fd = _design(6 * up, radial)
        rng = np.random.RandomState(up + 10 * radial)
        x = cu((rng.randn(2, 5, 40, 44) * 3).astype(np.float32), True)
        b = cu(rng.randn(5).astype(np.float32), True)
        pad = [11, 10, 11, 10] if up == 2 else [-2, -5, -2, -5]
        y = ops.filtered_lrelu.filtered_lrelu(x, cu(fu), cu(fd), b, up=up, down=2, padding=pad, clamp=6.0)
        dy = torch.randn_like(y)
        before = ops.capi.lib().sg3_launch_count()
        dx, db = torch.autograd.grad(y, [x, b], dy, retain_graph=True)
        assert ops.capi.lib().sg3_launch_count() - before == 1            # one fused launch, no separate reduction kernel of ours
        dx2, db2 = torch.autograd.grad(y, [x, b], dy, create_graph=True)   # reference-style path: db = dx.sum([0, 2, 3])
        assert torch.equal(dx, dx2.detach())
        assert rel_err(db.cpu().numpy(), db2.detach().cpu().numpy()) < 1e-5


@pytest.mark.parametrize('seed', list(range(24)))
def test_fused_random_shapes_forward_backward(ops, seed):
    """Randomised sweep over the fused kernels (strip / chunk edges, paddings of both signs, odd sizes, all filter kinds):
    forward, dx and db against the oracle."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(1000 + seed)
    up = int(rng.choice([2, 4]))
    radial = bool(rng.randint(2))
    fu, fd = _design(6 * up, radial)
    N, C = int(rng.randint(1, 3)), int(rng.randint(1, 5))
    H, W = int(rng.randint(20, 140)), int(rng.randint(20, 200))
    base = [11, 10, 11, 10] if up == 2 else [-2, -5, -2, -5]
    pad = [int(b + rng.randint(-6, 7)) for b in base]
    x = (rng.randn(N, C, H, W) * 3).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    kw = dict(up=up, down=2, padding=pad, gain=float(np.sqrt(2)), slope=0.2, clamp=float(rng.choice([4.0, 256.0])))
    try:
        y_ref, signs = orc.filtered_lrelu(x, fu, fd, b, return_signs=True, **kw)
    except Exception:
        pytest.skip('degenerate output size')
    xt, bt = cu(x, True), cu(b, True)
    y = ops.filtered_lrelu.filtered_lrelu(xt, cu(fu), cu(fd), bt, **kw)
    assert y.shape == y_ref.shape
    assert rel_err(y.detach().cpu().numpy(), y_ref) < TOL32
    dy = rng.randn(*y_ref.shape).astype(np.float32)
    kwb = dict(kw)
    kwb.pop('clamp')
    dx_ref, db_ref = orc.filtered_lrelu_bwd(dy, signs, x.shape, fu, fd, **kwb)
    dx, db = torch.autograd.grad(y, [xt, bt], cu(dy))
    assert rel_err(dx.cpu().numpy(), dx_ref) < 5e-5
    assert rel_err(db.cpu().numpy(), db_ref) < 5e-5
