"""Regression tests for the round-1 advisor findings that need a device."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pkg():
    import sg3_b200
    from sg3_b200 import modulated_conv, networks  # noqa: F401
    sg3_b200.filtered_lrelu._quiet_fallback = True
    return sg3_b200


def cu(a, dev='cuda'):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_missing_up_filter_with_factor_2_does_not_crash(pkg):
    """fu=None with up=2 / down=2 reaches the dense-up kernel with a NULL tap pointer (host-side segfault in round 1):
    NULL is the 1x1 identity filter.  Forward without grad, and the backward of a layer whose fd is None."""
    g = golden('ops.npz')
    c = g.case('flrelu', 'fuNone_up2_dn2')
    with torch.no_grad():
        y = pkg.filtered_lrelu.filtered_lrelu(cu(c['x']), None, cu(c['fd']), cu(c['b']), up=2, down=2,
                                              padding=[int(v) for v in c['padding']], gain=float(c['gain']),
                                              slope=float(c['slope']), clamp=float(c['clamp']))
    assert rel_err(y.cpu().numpy(), c['y']) < 2e-5
    c = g.case('flrelu', 'up2_dn2_fdNone')
    x, b = cu(c['x']).requires_grad_(True), cu(c['b']).requires_grad_(True)
    y = pkg.filtered_lrelu.filtered_lrelu(x, cu(c['fu']), None, b, up=2, down=2, padding=[int(v) for v in c['padding']],
                                          gain=float(c['gain']), slope=float(c['slope']), clamp=float(c['clamp']))
    assert rel_err(y.detach().cpu().numpy(), c['y']) < 2e-5
    dx, db = torch.autograd.grad(y, [x, b], cu(c['dy']))
    assert rel_err(dx.cpu().numpy(), c['dx']) < 2e-5 and rel_err(db.cpu().numpy(), c['db']) < 2e-5


def test_generator_moved_under_inference_mode(pkg):
    """Filter buffers created / moved under torch.inference_mode() have no version counter; the host-tap cache must cope."""
    torch.manual_seed(0)
    G = pkg.networks.Generator(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048,
                               channel_max=32, conv_kernel=1, use_radial_filters=True).eval().requires_grad_(False)
    z = torch.randn(2, 64, generator=torch.Generator().manual_seed(1))
    Gc = pkg.networks.Generator(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048,
                                channel_max=32, conv_kernel=1, use_radial_filters=True).eval().requires_grad_(False)
    Gc.load_state_dict(G.state_dict())
    with torch.no_grad():
        ref = Gc.cuda()(z.cuda(), None, noise_mode='const', force_fp32=True)
    with torch.inference_mode():
        G = G.to('cuda')
        img = G(z.cuda(), None, noise_mode='const', force_fp32=True)
        img2 = G(z.cuda(), None, noise_mode='const', force_fp32=True)
    assert torch.equal(img, img2)
    assert rel_err(img.cpu().numpy(), ref.cpu().numpy()) < 1e-6


def test_modulated_conv_unaligned_view_falls_back(pkg):
    """A contiguous view at an odd storage offset is not TMA-addressable (16-byte base): the tensor-core launch answers
    NOKERNEL and the op reruns the exact SIMT contraction instead of raising."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(5)
    N, I, O, H = 2, 32, 64, 20
    buf = torch.zeros(N * I * H * H + 1, device='cuda')
    x_np = rng.randn(N, I, H, H).astype(np.float32)
    x = buf[1:].view(N, I, H, H)
    x.copy_(cu(x_np))
    assert x.is_contiguous() and x.data_ptr() % 16 == 4
    w = rng.randn(O, I, 1, 1).astype(np.float32)
    s = (rng.randn(N, I) + 1).astype(np.float32)
    y = pkg.modulated_conv.modulated_conv2d(x, cu(w), cu(s), math='tf32')
    ref = orc.modulated_conv2d(x_np, w, s)
    assert rel_err(y.cpu().numpy(), ref) < 2e-3
    xh = buf[1:].view(N, I, H, H).half()
    hb = torch.zeros(N * I * H * H + 1, device='cuda', dtype=torch.float16)
    xh2 = hb[1:].view(N, I, H, H)
    xh2.copy_(xh)
    y16 = pkg.modulated_conv.modulated_conv2d(xh2, cu(w), cu(s), math='tf32')
    assert y16.dtype == torch.float16 and rel_err(y16.float().cpu().numpy(), ref) < 5e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two devices in one process')
def test_second_device_in_one_process(pkg):
    """cudaFuncSetAttribute is per device: kernels with > 48 KB of dynamic shared memory must also launch on cuda:1
    after cuda:0 was used first (process-wide call_once in round 1)."""
    torch.manual_seed(0)
    kw = dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32)
    z = torch.randn(2, 64, generator=torch.Generator().manual_seed(1))
    for extra in (dict(conv_kernel=1, use_radial_filters=True), dict(conv_kernel=3, use_radial_filters=False)):
        G = pkg.networks.Generator(**kw, **extra).eval()
        imgs = []
        for dev in ('cuda:0', 'cuda:1'):
            Gd = G.to(dev)
            ws = Gd.mapping(z.to(dev), None).detach().requires_grad_(True)
            img = Gd.synthesis(ws, noise_mode='const', force_fp32=True)
            img.square().mean().backward()
            torch.cuda.synchronize(dev)
            imgs.append(img.detach().cpu())
        assert torch.equal(imgs[0], imgs[1])
