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
