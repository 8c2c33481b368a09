"""Guard-band checks of every kernel family (the memcheck substitute: compute-sanitizer is closed on the GPU pool).

Every output / scratch tensor the op layer allocates (`torch.empty`, `torch.zeros`, `torch.empty_like` in the package's
Python host code) is carved out of a larger byte buffer pre-filled with 0xFF: 64 KiB of canary on either side and
NaN bit patterns (0xFFFFFFFF fp32, 0xFFFF fp16) in the body.  After the op:
  * both canaries must be untouched   -> no kernel stored outside the tensor it was given;
  * float outputs must hold no NaN    -> every element of every output was written (the inputs are finite).
Values are checked against the oracle elsewhere (test_ops_gpu.py, test_network_gpu.py); this file only looks at where
the kernels write, over shapes that stress strip / chunk / tile edges (odd sizes, ragged channel counts, both dtypes).
"""
import contextlib

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

GUARD = 1 << 16


class _Arena:
    def __init__(self):
        self.bufs = []          # (raw uint8 buffer, body bytes)
        self.bodies = []        # float views handed out by torch.empty (must be fully written by the op)
        self._orig = {}

    def _carve(self, shape, dtype, device, zero):
        shape = tuple(int(v) for v in shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
        raw = self._orig['empty'](2 * GUARD + ((nbytes + 255) // 256) * 256, dtype=torch.uint8, device=device)
        raw.fill_(0xFF)
        body = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if zero:
            body.zero_()
        elif dtype in (torch.float16, torch.float32):
            self.bodies.append(body)
        self.bufs.append((raw, nbytes))
        return body

    @staticmethod
    def _is_cuda(device):
        return device is not None and torch.device(device).type == 'cuda'

    def _alloc(self, name, zero):
        orig = self._orig[name]

        def fn(*size, **kw):
            if not self._is_cuda(kw.get('device')) or set(kw) - {'dtype', 'device'}:
                return orig(*size, **kw)
            shape = size[0] if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)) else size
            return self._carve(shape, kw.get('dtype', torch.float32), kw['device'], zero)
        return fn

    def _empty_like(self, x, **kw):
        if not x.is_cuda or kw:
            return self._orig['empty_like'](x, **kw)
        return self._carve(x.shape, x.dtype, x.device, False)

    def __enter__(self):
        self._orig = dict(empty=torch.empty, zeros=torch.zeros, empty_like=torch.empty_like)
        torch.empty = self._alloc('empty', False)
        torch.zeros = self._alloc('zeros', True)
        torch.empty_like = self._empty_like
        return self

    def __exit__(self, *exc):
        torch.empty, torch.zeros, torch.empty_like = self._orig['empty'], self._orig['zeros'], self._orig['empty_like']
        return False

    def check(self, min_bufs=1, scratch_ok=False):
        torch.cuda.synchronize()
        assert len(self.bufs) >= min_bufs, f'only {len(self.bufs)} guarded allocations: the patch missed the op'
        for raw, nbytes in self.bufs:
            end = GUARD + nbytes
            assert bool((raw[:GUARD] == 0xFF).all()), f'store below a {nbytes}-byte tensor'
            assert bool((raw[end:] == 0xFF).all()), f'store above a {nbytes}-byte tensor'
        if not scratch_ok:
            for body in self.bodies:
                assert not bool(torch.isnan(body).any()), f'output {tuple(body.shape)} {body.dtype} not fully written'


@contextlib.contextmanager
def arena():
    a = _Arena()
    with a:
        yield a


@pytest.fixture(scope='module')
def pkg():
    import sg3_b200
    from sg3_b200 import modulated_conv, networks  # noqa: F401
    assert torch.cuda.is_available()
    sg3_b200.filtered_lrelu._quiet_fallback = True
    return sg3_b200


def _design(pkg, taps, radial):
    from oracle import sg3_oracle as orc                 # filter design only (taps are data)
    _, specs = orc.layer_specs(64, channel_base=2048, channel_max=32, conv_kernel=1, use_radial_filters=radial)
    sp = next(s for s in specs if s['up'] * 6 == taps and np.asarray(s['down_filter']).size > 1)
    return torch.from_numpy(np.asarray(sp['up_filter'])).cuda(), torch.from_numpy(np.asarray(sp['down_filter'])).cuda()


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('seed', list(range(12)))
def test_guard_fused_filtered_lrelu(pkg, seed, dtype):
    """Fused forward (sign-free and sign WRITE) and fused backward (sign READ, bias-gradient atomics)."""
    rng = np.random.RandomState(7000 + seed)
    up = 2 if seed % 2 == 0 else 4
    radial = bool((seed // 2) % 2)
    fu, fd = _design(pkg, 6 * up, radial)
    N, C = int(rng.randint(1, 3)), int(rng.randint(1, 6))
    H, W = int(rng.randint(20, 150)), int(rng.randint(20, 260))
    base = [11, 10, 11, 10] if up == 2 else [-2, -5, -2, -5]
    pad = [int(b + rng.randint(-5, 6)) for b in base]
    x = (torch.randn(N, C, H, W, device='cuda') * 3).to(dtype)
    b = torch.randn(C, device='cuda').to(dtype)
    kw = dict(up=up, down=2, padding=pad, clamp=4.0)
    launches = pkg.capi.lib().sg3_launch_count()
    with arena() as a:
        y0 = pkg.filtered_lrelu.filtered_lrelu(x, fu, fd, b, **kw)                       # no signs
        xg, bg = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y1 = pkg.filtered_lrelu.filtered_lrelu(xg, fu, fd, bg, **kw)                     # sign WRITE
        dx, db = torch.autograd.grad(y1, [xg, bg], torch.randn_like(y1))                 # sign READ + ysum
        a.check(min_bufs=5)                                                              # y, y, signs, ysum, dx
    assert pkg.capi.lib().sg3_launch_count() - launches == 3, 'expected three fused launches (no generic composition)'
    assert torch.equal(y0, y1.detach())
    assert dx.shape == x.shape and bool(torch.isfinite(dx).all()) and bool(torch.isfinite(db).all())


@pytest.mark.parametrize('case', ['generic_up3', 'fp64', 'pointwise', 'down4_sep'])
def test_guard_other_filtered_lrelu_paths(pkg, case):
    """The generic composition (bias_act -> upfirdn2d -> act+signs -> upfirdn2d), the one-pass ToRGB kernel and the
    separable down-4 backward shape."""
    rng = np.random.RandomState(3)
    dtype = torch.float64 if case == 'fp64' else torch.float32
    x = (torch.randn(2, 3, 37, 45, device='cuda') * 3).to(dtype).requires_grad_(True)
    b = torch.randn(3, device='cuda').to(dtype).requires_grad_(True)
    if case == 'pointwise':
        kw = dict(fu=None, fd=None, up=1, down=1, padding=0, clamp=2.0)
    elif case == 'generic_up3':
        f = torch.from_numpy(rng.rand(9).astype(np.float32)).cuda()
        kw = dict(fu=f, fd=f, up=3, down=2, padding=[4, 3, 4, 3], clamp=2.0)
    elif case == 'down4_sep':
        fu, _ = _design(pkg, 24, False)
        fd12, _ = _design(pkg, 12, False)
        kw = dict(fu=fd12, fd=fu, up=2, down=4, padding=[16, 17, 16, 17], clamp=None)
    else:
        fu, fd = _design(pkg, 12, True)
        kw = dict(fu=fu, fd=fd, up=2, down=2, padding=[11, 10, 11, 10], clamp=2.0)
    with arena() as a:
        y = pkg.filtered_lrelu.filtered_lrelu(x, b=b, **kw)
        dx, db = torch.autograd.grad(y, [x, b], torch.randn_like(y))
        # The host allocates an output before the library answers SG3_E_NOKERNEL (sign WRITE on the dense-up / pointwise
        # kernels, ysum on the pointwise kernel, upfirdn2d_sep for up = 3); such a buffer is dropped unwritten and the
        # generic composition allocates its own -- so bodies are not NaN-checked here, the results below are.
        a.check(min_bufs=2, scratch_ok=True)
    assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(dx).all()) and bool(torch.isfinite(db).all())


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_guard_upfirdn2d_and_bias_act(pkg, dtype):
    rng = np.random.RandomState(5)
    f12 = torch.from_numpy(rng.rand(12).astype(np.float32)).cuda()
    f24 = torch.from_numpy(rng.rand(24).astype(np.float32)).cuda()
    f4 = torch.from_numpy(rng.rand(4, 4).astype(np.float32)).cuda()
    f7 = torch.from_numpy(rng.rand(5, 7).astype(np.float32)).cuda()
    x = torch.randn(2, 3, 61, 83, device='cuda').to(dtype)
    with arena() as a:
        for f, up, down, pad in ((f12, 2, 1, [6, 5, 6, 5]), (f24, 4, 1, [13, 10, 13, 10]), (f12, 1, 2, [5, 5, 5, 5]),
                                 (f24, 1, 4, [10, 10, 10, 10]), (f12, 2, 1, [-3, -4, 2, -7]), (f4, 1, 1, [1, 2, 1, 2]),
                                 (f4, 2, 1, [2, 1, 2, 1]), (f4, 1, 2, [1, 1, 1, 1]), (f7, 3, 2, [3, 1, 0, 4]),
                                 (f7, [1, 3], [2, 1], [0, 5, 2, 0]), (None, 2, 1, 0), (f12, [2, 1], [1, 2], [3, 3, 4, 4])):
            y = pkg.upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=pad, gain=1.5)
            assert y.numel() > 0
        xc = x.contiguous(memory_format=torch.channels_last)
        pkg.upfirdn2d.upfirdn2d(xc, f12, up=2, padding=[6, 5, 6, 5])
        a.check(min_bufs=12)
    xb = torch.randn(5, 7, 33, 29, device='cuda').to(dtype).requires_grad_(True)
    bb = torch.randn(7, device='cuda').to(dtype).requires_grad_(True)
    with arena() as a:
        for act in pkg.bias_act.activation_funcs:
            y = pkg.bias_act.bias_act(xb, bb, act=act, clamp=1.5)
            dx, db = torch.autograd.grad(y, [xb, bb], torch.randn_like(y), create_graph=True)
            if dx.requires_grad:
                torch.autograd.grad(dx.sum(), xb, allow_unused=True)
        y = pkg.bias_act.bias_act(torch.randn(3, 517, device='cuda').to(dtype), torch.randn(517, device='cuda').to(dtype), act='lrelu')
        a.check(min_bufs=10)


@pytest.mark.parametrize('seed', list(range(10)))
def test_guard_modulated_conv2d(pkg, seed):
    """Weight prologue, exact-FP32 SIMT conv, 1x1 / 3x3 TF32 tcgen05 contractions, native dgrad / wgrad and the fused chain rule."""
    rng = np.random.RandomState(9000 + seed)
    k = 1 if seed % 2 == 0 else 3
    N, I, O = int(rng.randint(1, 4)), int(rng.randint(3, 200)), int(rng.randint(3, 300))
    H, W = int(rng.randint(5, 90)), 4 * int(rng.randint(2, 60))
    if k == 1 and (H * W) % 4:
        H += 4 - (H % 4)
    x = torch.randn(N, I, H, W, device='cuda', requires_grad=True)
    w = torch.randn(O, I, k, k, device='cuda', requires_grad=True)
    s = torch.randn(N, I, device='cuda', requires_grad=True)
    g = torch.tensor(0.8, device='cuda')
    for math in ('tf32', 'fp32'):
        with arena() as a:
            y = pkg.modulated_conv.modulated_conv2d(x, w, s, demodulate=True, padding=k - 1, input_gain=g, math=math)
            grads = torch.autograd.grad(y, [x, w, s], torch.randn_like(y))
            # weight buffers carry zero-filled / unwritten pitch padding by design (ldw = ceil32): bodies are not NaN-checked
            a.check(min_bufs=2, scratch_ok=True)
        assert y.shape == (N, O, H + k - 1, W + k - 1) and bool(torch.isfinite(y).all())
        assert all(bool(torch.isfinite(t).all()) for t in grads)
    if k == 1:
        with arena() as a:                       # fp16 layers: kind::f16 MMAs on fp16 activations and weights
            y = pkg.modulated_conv.modulated_conv2d(x.detach().half(), w.detach(), s.detach(), demodulate=True, padding=0, input_gain=g)
            a.check(min_bufs=2, scratch_ok=True)
        assert bool(torch.isfinite(y).all())


TINY_CFG = dict(
    tinyR=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=1, use_radial_filters=True),
    tinyT=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=3, use_radial_filters=False),
)


@pytest.mark.parametrize('force_fp32', [True, False])
@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_guard_tiny_generator_forward_backward(pkg, name, force_fp32):
    """A whole 15-layer generator, forward and backward (PTI-style), with every op-layer allocation guarded."""
    torch.manual_seed(0)
    G = pkg.networks.Generator(**TINY_CFG[name]).cuda().eval().requires_grad_(True)
    ws = torch.from_numpy(golden('tiny.npz').z[f'{name}/ws']).cuda()
    pkg.modulated_conv.set_math('tf32')
    try:
        with arena() as a:
            img = G.synthesis(ws, noise_mode='const', force_fp32=force_fp32)
            loss = img.float().square().mean()
            grads = torch.autograd.grad(loss, list(G.synthesis.parameters()), allow_unused=True)
            a.check(min_bufs=45, scratch_ok=True)
    finally:
        pkg.modulated_conv.set_math(None)
    assert bool(torch.isfinite(img).all())
    assert all(g is None or bool(torch.isfinite(g).all()) for g in grads)


@pytest.mark.parametrize('case', ['R_up2_dense', 'R_up4_dense', 'T_up2_sep'])
def test_fused_kernels_are_deterministic(pkg, case):
    """The racecheck substitute: warps of the stream kernels share nothing and order their private shared-memory ring
    with __syncwarp / mbarriers only, so a missing fence shows up as run-to-run differences.  The same launch is repeated
    with the SMs oversubscribed (several waves, a second stream competing) and must be bit-identical every time: forward,
    sign tensor and dx.  (db uses fp32 atomics and is order-dependent by design; it is not compared.)"""
    up, radial = {'R_up2_dense': (2, True), 'R_up4_dense': (4, True), 'T_up2_sep': (2, False)}[case]
    fu, fd = _design(pkg, 6 * up, radial)
    torch.manual_seed(1)
    x = (torch.randn(2, 24, 276 if up == 2 else 148, 276 if up == 2 else 148, device='cuda') * 3).requires_grad_(True)
    b = torch.randn(24, device='cuda')
    pad = [11, 10, 11, 10] if up == 2 else [-2, -5, -2, -5]
    side = torch.cuda.Stream()
    noise = torch.randn(1 << 24, device='cuda')
    ref = None
    for it in range(6):
        with torch.cuda.stream(side):                      # a competing memory-bound kernel on another stream
            noise.mul_(1.0001)
        y = pkg.filtered_lrelu.filtered_lrelu(x, fu, fd, b, up=up, down=2, padding=pad, clamp=4.0)
        nb = (2 * y.shape[3] - 1 + 11) // 4                # whole bytes of the active sign region (the 16-pixel row padding is never written)
        signs = y.grad_fn.saved_tensors[0][..., :nb].clone()      # the sign tensor the forward kernel wrote (before backward frees it)
        (dx,) = torch.autograd.grad(y, x, torch.ones_like(y))
        cur = (y.detach().clone(), dx.clone(), signs)
        if ref is None:
            ref = cur
        else:
            assert torch.equal(cur[0], ref[0]), f'forward differs in run {it}'
            assert torch.equal(cur[1], ref[1]), f'dx differs in run {it}'
            assert torch.equal(cur[2], ref[2]), f'sign tensor differs in run {it}'
    torch.cuda.synchronize()


def test_tensor_core_conv_is_deterministic(pkg):
    """Same for the tcgen05 kernels (TMA ring, TMEM double buffering, mbarrier hand-overs): forward and dgrad of the 1x1 and
    3x3 contractions are bit-identical from run to run (wgrad is split-K with fp32 atomics: order-dependent by design)."""
    torch.manual_seed(2)
    for k, (I, O, H, W) in ((1, (161, 102, 96, 128)), (3, (81, 51, 94, 132)), (3, (40, 32, 70, 224))):
        x = torch.randn(3, I, H, W, device='cuda', requires_grad=True)
        w = torch.randn(O, I, k, k, device='cuda')
        s = torch.randn(3, I, device='cuda')
        ref = None
        for it in range(5):
            y = pkg.modulated_conv.modulated_conv2d(x, w, s, demodulate=True, padding=k - 1, math='tf32')
            (dx,) = torch.autograd.grad(y, x, torch.ones_like(y))
            if ref is None:
                ref = (y.detach().clone(), dx.clone())
            else:
                assert torch.equal(y.detach(), ref[0]), f'k={k} forward differs in run {it}'
                if k == 1:                              # 3x3 dgrad is cuDNN (may pick non-deterministic algorithms)
                    assert torch.equal(dx, ref[1]), f'k={k} dgrad differs in run {it}'
