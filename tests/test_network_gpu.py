"""GPU parity of modulated_conv2d and of whole generators against the golden vectors generated
from the reference (tests/golden/make_golden.py).

Tolerances (max |err| / max |ref|): fp32 math 1e-4 for a 15-layer network (north_star budget
1e-3); the TF32 tensor-core contraction is stated separately: 2e-3 per conv, 1e-2 per network.
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

TINY_CFG = dict(
    tinyR=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=1, use_radial_filters=True),
    tinyT=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=3, use_radial_filters=False),
)


@pytest.fixture(scope='module')
def pkg():
    import sg3_b200
    from sg3_b200 import modulated_conv, networks  # noqa: F401
    sg3_b200.filtered_lrelu._quiet_fallback = True
    return sg3_b200


def cu(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().requires_grad_(grad)


@pytest.mark.parametrize('math,tol', [('fp32', 2e-5), ('tf32', 2e-3)])
@pytest.mark.parametrize('case', golden('modconv.npz').cases('modconv'))
def test_modulated_conv2d_golden(pkg, case, math, tol):
    c = golden('modconv.npz').case('modconv', case)
    k = c['w'].shape[-1]
    x, w, s = cu(c['x'], True), cu(c['w'], True), cu(c['s'], True)
    g = cu(c['input_gain']) if 'input_gain' in c else None
    y = pkg.modulated_conv.modulated_conv2d(x, w, s, demodulate=bool(c['demodulate']), padding=k - 1, input_gain=g, math=math)
    assert tuple(y.shape) == c['y'].shape
    assert rel_err(y.detach().cpu().numpy(), c['y']) < tol
    dx, dw, ds = torch.autograd.grad(y, [x, w, s], cu(c['dy']))
    assert rel_err(dx.cpu().numpy(), c['dx']) < 2e-3
    assert rel_err(dw.cpu().numpy(), c['dw']) < 2e-3
    assert rel_err(ds.cpu().numpy(), c['ds']) < 2e-3


def test_modconv_weights_vs_oracle(pkg):
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(3)
    w = rng.randn(37, 29, 3, 3).astype(np.float32)
    s = (rng.randn(5, 29) + 1).astype(np.float32)
    ref = orc.modconv_weights(w, s, True, np.float32(0.7)).reshape(5, 37, -1)
    got = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda())
    assert rel_err(got[:, :, :29 * 9].cpu().numpy(), ref) < 1e-5 and float(got[:, :, 29 * 9:].abs().max()) == 0
    ref = orc.modconv_weights(w, s, False, None).reshape(5, 37, -1)
    got = pkg.modulated_conv.modconv_weights(cu(w), cu(s), False, None)
    assert rel_err(got[:, :, :29 * 9].cpu().numpy(), ref) < 1e-6


def _build(pkg, name):
    torch.manual_seed(0)
    G = pkg.networks.Generator(**TINY_CFG[name]).eval().requires_grad_(False)
    g = golden('tiny.npz')
    sd = {k: torch.from_numpy(v) for k, v in g.sub(f'{name}/state/').items()}
    for k, v in G.synthesis.state_dict().items():     # same seed -> same random init as the reference
        assert torch.equal(v, sd[k]), k
    return G.cuda(), g


@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_tiny_generator_fp32(pkg, name):
    G, g = _build(pkg, name)
    pkg.modulated_conv.set_math('fp32')
    try:
        z = cu(g.z[f'{name}/z'])
        ws = G.mapping(z, None)
        assert rel_err(ws.cpu().numpy(), g.z[f'{name}/ws']) < 1e-5          # mapping network (bias_act lrelu)
        ws = cu(g.z[f'{name}/ws'])
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
        assert rel_err(img.cpu().numpy(), g.z[f'{name}/img']) < 1e-4
        x = G.synthesis.input(ws[:, 0])
        assert rel_err(x[:1, :8].cpu().numpy(), g.z[f'{name}/feat/input']) < 1e-5
        for i, lname in enumerate(G.synthesis.layer_names[:13]):
            x = getattr(G.synthesis, lname)(x, ws[:, i + 1], noise_mode='const', force_fp32=True)
            if i in (0, 2, 7, 12):
                assert rel_err(x[:1, :8].cpu().numpy(), g.z[f'{name}/feat/after_{i}']) < 1e-4, lname
        # S-space path gives the same image
        img2 = G.synthesis(None, all_s=G.synthesis.W2S(ws), noise_mode='const', force_fp32=True)
        assert rel_err(img2.cpu().numpy(), g.z[f'{name}/img']) < 1e-4
    finally:
        pkg.modulated_conv.set_math(None)


@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_tiny_generator_tf32(pkg, name):
    """TF32 tensor-core convs.  The tensor core truncates its fp32 activations; the three policies of
    modulated_conv.set_tf32_activation_policy: 'compensate' (default: the expected truncation loss is folded into the weights),
    'round' (the producing stencil rounds its outputs to the nearest TF32 value), 'truncate' (nothing).  Measured on B200: 3.1e-3 /
    2.3e-3 (tinyR / tinyT) compensated, 2.3e-3 / 3.3e-3 rounded, 7.1e-3 / 6.2e-3 truncated; the reference's cuDNN TF32 path: 3.2e-3 (tinyT)."""
    from sg3_b200 import modulated_conv as mc
    G, g = _build(pkg, name)
    mc.set_math('tf32')
    errs = {}
    try:
        for policy in ('compensate', 'round', 'truncate'):
            mc.set_tf32_activation_policy(policy)
            img = G.synthesis(cu(g.z[f'{name}/ws']), noise_mode='const', force_fp32=True)
            errs[policy] = rel_err(img.cpu().numpy(), g.z[f'{name}/img'])
        print(name, {k: f'{v:.2e}' for k, v in errs.items()})
        assert errs['compensate'] < 5e-3 and errs['round'] < 5e-3 and errs['truncate'] < 1e-2
        assert errs['compensate'] < errs['truncate'] and errs['round'] < errs['truncate']
        # the 'round' switch of the fused kernel: TF32-representable outputs within half a TF32 ulp of the unrounded ones
        x = torch.randn(2, 8, 20, 20, device='cuda')
        L = getattr(G.synthesis, G.synthesis.layer_names[3])
        fl = pkg.filtered_lrelu
        kw = dict(fu=L.up_filter, fd=L.down_filter, up=L.up_factor, down=L.down_factor, padding=L.padding, clamp=256)
        y0 = fl.filtered_lrelu(x, **kw)
        with fl.tf32_rounded_outputs(True):
            y1 = fl.filtered_lrelu(x, **kw)
        assert int((y1.view(torch.int32) & 0x1fff).abs().max()) == 0
        assert float((y1 - y0).abs().max()) <= float(y0.abs().max()) * 2.0 ** -11
        assert not torch.equal(y0, y1)
        # the 'compensate' switch of the weight prologue: weights scaled by 1 + 3.52e-4 before their own rounding
        w = torch.randn(24, 40, 1, 1, device='cuda')
        sty = torch.randn(2, 40, device='cuda')
        w1 = mc.modconv_weights(w, sty, round_tf32=True)
        w4 = mc.modconv_weights(w, sty, round_tf32=True, compensate=True)
        ratio = (w4[:, :, :40] / w1[:, :, :40]).double().mean().item()
        assert abs(ratio - (1 + 3.5217e-4)) < 1e-4
    finally:
        mc.set_tf32_activation_policy('compensate')
        mc.set_math(None)


@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_tiny_generator_fp16_layers(pkg, name):
    """force_fp32=False: layers flagged use_fp16 run with fp16 activations (reference :355)."""
    G, g = _build(pkg, name)
    img = G.synthesis(cu(g.z[f'{name}/ws']), noise_mode='const')
    assert img.dtype == torch.float32
    assert rel_err(img.cpu().numpy(), g.z[f'{name}/img']) < 2e-2


@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_tiny_generator_gradients(pkg, name):
    """PTI-style: gradients of a scalar loss wrt every synthesis parameter vs the reference's autograd."""
    G, g = _build(pkg, name)
    pkg.modulated_conv.set_math('fp32')
    try:
        G.requires_grad_(True)
        img = G.synthesis(cu(g.z[f'{name}/ws']), noise_mode='const', force_fp32=True)
        loss = (img * cu(g.z[f'{name}/tgt'])).mean()
        assert abs(float(loss) - float(g.z[f'{name}/loss'])) < 1e-5 * max(1.0, abs(float(g.z[f'{name}/loss']))) + 1e-7
        params = dict(G.synthesis.named_parameters())
        grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
        checked = 0
        for (k, _), gr in zip(params.items(), grads):
            key = f'{name}/grad/{k}'
            if key in g.z.files:
                ref = g.z[key]
                if np.abs(ref).max() > 0:
                    assert rel_err(gr.cpu().numpy(), ref) < 1e-3, k
                    checked += 1
        assert checked >= 55
    finally:
        pkg.modulated_conv.set_math(None)


@pytest.mark.parametrize('tag', ['per_sample', 'single'])
@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_user_transform(pkg, name, tag):
    """BASELINE configs[3] (ReStyle / pSp inversion) sets one landmarks transform per sample on `synthesis.input.transform`
    (models/setgan/encoder/psp3.py:63-65), the video FOV expansion translates it (utils/fov_expansion.py:20-27): [N, 3, 3] and
    [3, 3] user transforms vs the reference, eagerly and through the CUDA-graph wrapper (in-place update of the buffer)."""
    G, _ = _build(pkg, name)
    g = golden('transform.npz').z
    ws, m = cu(g[f'{name}/ws']), cu(g[f'{name}/{tag}/transform'])
    pkg.modulated_conv.set_math('fp32')
    try:
        G.synthesis.input.transform = m
        x = G.synthesis.input(ws[:, 0])
        assert rel_err(x[:, :8].cpu().numpy(), g[f'{name}/{tag}/input']) < 1e-5
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
        assert rel_err(img.cpu().numpy(), g[f'{name}/{tag}/img']) < 1e-4
        # graph captured with identity transforms of the same shape, then the buffer is updated in place and replayed
        G.synthesis.input.transform = torch.eye(3, device='cuda').expand_as(m).contiguous()
        graphed = pkg.networks.GraphedSynthesis(G.synthesis, ws)
        G.synthesis.input.transform.copy_(m)
        assert rel_err(graphed(ws).cpu().numpy(), g[f'{name}/{tag}/img']) < 1e-4
    finally:
        pkg.modulated_conv.set_math(None)


@pytest.mark.parametrize('margins', [(10, 6, 4, 12), (0, 8, 0, 0), (5, 0, 0, 7)])
def test_fov_expander_matches_sequential_views(pkg, margins):
    """sg3_b200.fov.Expander (one batched call with per-sample transforms) vs the reference's procedure restated here
    (utils/fov_expansion.py:14-31, 88-110: one batch-1 synthesis call per view with a single [3, 3] transform, then paste)."""
    from sg3_b200 import fov
    G, g = _build(pkg, 'tinyR')
    pkg.modulated_conv.set_math('fp32')      # exact contraction: with TF32 the batch-global style RMS moves weight roundings
    try:
        _fov_check(pkg, fov, G, g, margins)
    finally:
        pkg.modulated_conv.set_math(None)


def _fov_check(pkg, fov, G, g, margins):
    res = G.img_resolution
    right, left, top, bottom = margins
    ws = cu(g.z['tinyR/ws'])
    ang = 0.2
    lt = np.array([[np.cos(ang), np.sin(ang), 0.05], [-np.sin(ang), np.cos(ang), -0.02], [0, 0, 1]])
    views = fov.view_transforms(res, right, left, top, bottom)
    imgs = []
    for t in views:
        if t is None:
            imgs.append(None)
            continue
        G.synthesis.input.transform = torch.from_numpy(lt @ t).float().cuda()
        with torch.no_grad():
            imgs.append(G.synthesis(ws, noise_mode='const', force_fp32=True))
    want = torch.zeros(ws.shape[0], 3, top + res + bottom, left + res + right, device='cuda')
    want[:, :, top:top + res, left:left + res] = imgs[0]
    if left:
        want[:, :, top:top + res, :left] = imgs[1][:, :, :, :left]
    if top:
        want[:, :, :top, left:left + res] = imgs[2][:, :, :top, :]
    if right:
        want[:, :, top:top + res, left + res:] = imgs[3][:, :, :, res - right:]
    if bottom:
        want[:, :, top + res:, left:left + res] = imgs[4][:, :, res - bottom:, :]
    if top and left:
        want[:, :, :top, :left] = imgs[5][:, :, :top, :left]
    if top and right:
        want[:, :, :top, res + left:] = imgs[6][:, :, :top, res - right:]
    if bottom and right:
        want[:, :, res + top:, res + left:] = imgs[7][:, :, res - bottom:, res - right:]
    if bottom and left:
        want[:, :, res + top:, :left] = imgs[8][:, :, res - bottom:, :left]
    eye = torch.eye(3, device='cuda')
    G.synthesis.input.transform = eye
    ex = fov.Expander(G)
    got = ex.generate_expanded_image(ws=ws, landmark_t=lt, pixels_right=right, pixels_left=left, pixels_top=top,
                                     pixels_bottom=bottom, noise_mode='const', force_fp32=True)
    assert got.shape == want.shape and G.synthesis.input.transform is eye
    # per-sample kernels throughout; the only cross-sample term is the batch-global style RMS, which demodulation cancels
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5
    got_s = ex.generate_expanded_image(all_s=G.synthesis.W2S(ws), landmark_t=torch.from_numpy(lt), pixels_right=right,
                                       pixels_left=left, pixels_top=top, pixels_bottom=bottom, noise_mode='const', force_fp32=True)
    assert rel_err(got_s.cpu().numpy(), want.cpu().numpy()) < 1e-5


def test_r256_config1(pkg):
    """BASELINE.json configs[0]: StyleGAN3-R 256^2, seed-0 random init, batch 4, fp32 -- against the image the
    reference produced on CPU with impl='ref' (strided subsample stored in tests/golden/r256.npz)."""
    g = golden('r256.npz').z
    torch.manual_seed(0)
    G = pkg.networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=256, img_channels=3,
                               **pkg.networks.CONFIG_R).eval().requires_grad_(False)
    sd = G.synthesis.state_dict()
    names = G.synthesis.layer_names
    for key in ['input.weight', 'input.freqs', names[0] + '.weight', names[13] + '.weight', names[14] + '.affine.weight']:
        assert np.array_equal(sd[key].flatten()[:64].numpy(), g['r256/probe/' + key]), key
    G = G.cuda()
    pkg.modulated_conv.set_math('fp32')
    try:
        img = G.synthesis(cu(g['r256/ws']), noise_mode='const', force_fp32=True)
    finally:
        pkg.modulated_conv.set_math(None)
    sub = img[:, :, ::4, ::4].cpu().numpy()
    assert rel_err(sub, g['r256/img_sub4']) < 1e-3
    stats = np.asarray([float(img.mean()), float(img.std()), float(img.abs().max())])
    assert np.allclose(stats, g['r256/stats'], rtol=1e-3, atol=1e-5)


# ---------------------------------------------------------------------------------------------
# TF32 tcgen05/TMEM contraction (modconv_tc.cu), called through the C ABI.

TC_CASES = [
    # N, I, O, H, W
    (1, 32, 16, 16, 16),        # single k-tile, single tile
    (2, 64, 64, 36, 36),        # P = 1296 (ragged last pixel tile)
    (1, 161, 102, 40, 44),      # ragged I and O (R-1024 L11 channel counts)
    (2, 256, 161, 52, 52),      # O -> 176-wide tile, 8 k-tiles (pipeline wraps twice)
    (1, 1024, 1024, 36, 36),    # four 256-wide n-tiles, 32 k-tiles
    (1, 645, 406, 20, 36),      # 2 n-tiles of 208
    (3, 64, 3, 64, 64),         # ToRGB: O = 3 -> 16-wide tile
]


@pytest.mark.parametrize('shape', TC_CASES, ids=['x'.join(map(str, s)) for s in TC_CASES])
def test_modconv_tc_vs_oracle(pkg, shape):
    """Tolerance for TF32 operands (10-bit mantissa, fp32 accumulate), stated separately from fp32 parity:
    max |err| <= 2e-3 * max |ref|."""
    from oracle import sg3_oracle as orc
    from sg3_b200 import capi
    N, I, O, H, W = shape
    rng = np.random.RandomState(I + O)
    x = rng.randn(N, I, H, W).astype(np.float32)
    wm = (rng.randn(N, O, I) / np.sqrt(I)).astype(np.float32)
    ldw = (I + 31) // 32 * 32
    wpad = np.zeros((N, O, ldw), np.float32)
    wpad[:, :, :I] = wm
    xt, wt = cu(x), cu(wpad)
    y = torch.empty(N, O, H, W, device='cuda')
    rc = capi.lib().sg3_modconv_fwd(xt.data_ptr(), wt.data_ptr(), y.data_ptr(), N, I, O, H, W, 1, 0, ldw, 1, capi.SG3_F32,
                                    capi.stream_ptr(xt.device))
    assert rc == 0, 'tensor-core path must take 1x1 convolutions'
    torch.cuda.synchronize()
    ref = orc.conv2d(x, wm.reshape(N, O, I, 1, 1), padding=0)
    assert rel_err(y.cpu().numpy(), ref) < 2e-3
    # and the exact path on the same padded weights agrees with the oracle to fp32 rounding
    y2 = torch.empty_like(y)
    rc = capi.lib().sg3_modconv_fwd(xt.data_ptr(), wt.data_ptr(), y2.data_ptr(), N, I, O, H, W, 1, 0, ldw, 0, capi.SG3_F32,
                                    capi.stream_ptr(xt.device))
    assert rc == 0
    assert rel_err(y2.cpu().numpy(), ref) < 1e-5


# ---------------------------------------------------------------------------------------------
# 3x3 TF32 tcgen05/TMEM contraction (modconv_tc3.cu, config T), called through the C ABI.

TC3_CASES = [
    # N, I, O, H, W, pad
    (1, 32, 16, 8, 8, 2),          # one tile, one chunk, O << 128
    (1, 32, 128, 36, 36, 2),       # T-1024 L0 map size, full 128-lane tile
    (2, 64, 64, 20, 52, 2),        # two chunks, two samples, non-square
    (1, 81, 51, 40, 148, 2),       # ragged I and O (T-1024 L11 channel counts), several column tiles
    (1, 512, 323, 16, 24, 2),      # 16 chunks, three channel tiles (ragged last)
    (1, 40, 200, 12, 276, 2),      # wide rows: several column tiles per row
    (2, 48, 32, 18, 20, 0),        # padding 0 ('valid'), the dgrad form
    (1, 128, 81, 30, 532, 2),      # short K loop, 532-wide rows
    (2, 96, 128, 96, 128, 2),      # streamed weights (3 chunks x 9 taps do not fit the ring), full 128-channel block
    (1, 64, 256, 75, 100, 2),      # odd number of pixel tiles, two channel blocks
    (3, 512, 512, 36, 36, 2),      # T-1024 L0 at batch 3: 16 chunks, four channel blocks
    (1, 51, 32, 40, 276, 2),       # O = 32: the three kx taps stacked along M (T-1024 L12 channel counts), several column tiles
    (2, 32, 32, 70, 148, 2),       # O = 32, one chunk (T-1024 L13), ragged last row block
    (1, 160, 32, 33, 532, 2),      # O = 32, five chunks: the stacked weights are streamed, not resident
]


@pytest.mark.parametrize('shape', TC3_CASES, ids=['x'.join(map(str, s)) for s in TC3_CASES])
def test_modconv_tc3_vs_oracle(pkg, shape):
    """3x3 tensor-core contraction vs the CPU oracle; TF32 tolerance (stated separately from fp32 parity):
    max |err| <= 2e-3 * max |ref|.  The exact SIMT kernel on the same weights must agree to fp32 rounding."""
    from oracle import sg3_oracle as orc
    from sg3_b200 import capi
    N, I, O, H, W, pad = shape
    rng = np.random.RandomState(I + O + W)
    x = rng.randn(N, I, H, W).astype(np.float32)
    wm = (rng.randn(N, O, I, 3, 3) / np.sqrt(9 * I)).astype(np.float32)
    assert capi.lib().sg3_modconv_tc_supported(I, O, H, W, 3, pad) == 0
    ldw = (I + 31) // 32 * 32
    wtap = np.zeros((N, 9, O, ldw), np.float32)                        # tap-major operand layout
    wtap[:, :, :, :I] = wm.reshape(N, O, I, 9).transpose(0, 3, 1, 2)
    OH, OW = H + 2 * pad - 2, W + 2 * pad - 2
    xt, wt = cu(x), cu(wtap)
    y = torch.full((N, O, OH, OW), float('nan'), device='cuda')
    rc = capi.lib().sg3_modconv_fwd(xt.data_ptr(), wt.data_ptr(), y.data_ptr(), N, I, O, H, W, 3, pad, ldw, 1, capi.SG3_F32,
                                    capi.stream_ptr(xt.device))
    assert rc == 0, 'tensor-core path must take 3x3 convolutions with padding 0 / 2'
    torch.cuda.synchronize()
    ref = orc.conv2d(x, wm, padding=pad)
    assert rel_err(y.cpu().numpy(), ref) < 2e-3
    ld0 = (9 * I + 31) // 32 * 32
    w0 = np.zeros((N, O, ld0), np.float32)
    w0[:, :, :9 * I] = wm.reshape(N, O, 9 * I)
    y2 = torch.empty_like(y)
    w0t = cu(w0)
    rc = capi.lib().sg3_modconv_fwd(xt.data_ptr(), w0t.data_ptr(), y2.data_ptr(), N, I, O, H, W, 3, pad, ld0, 0, capi.SG3_F32,
                                    capi.stream_ptr(xt.device))
    assert rc == 0
    assert rel_err(y2.cpu().numpy(), ref) < 1e-5


def test_modconv_tap_major_weights(pkg):
    """The tap-major prologue layout holds the same numbers as the standard layout."""
    rng = np.random.RandomState(5)
    w, s = rng.randn(37, 19, 3, 3).astype(np.float32), rng.randn(3, 19).astype(np.float32)
    a = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda())[:, :, :19 * 9].reshape(3, 37, 19, 9)
    b = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda(), tap_major=True)[:, :, :, :19]
    assert torch.equal(a.permute(0, 3, 1, 2), b)


def test_host_pipeline_matches_direct_call(pkg):
    """sharding.HostPipeline (pinned host in/out, copies overlapped with the next forward) returns the same images."""
    from sg3_b200 import sharding
    G, g = _build(pkg, 'tinyR')
    ws = cu(g.z['tinyR/ws'])
    ref = G.synthesis(ws, noise_mode='const', force_fp32=True).cpu()
    ws_host = ws.cpu().pin_memory()
    outs = [torch.empty_like(ref).pin_memory() for _ in range(2)]
    pipe = sharding.HostPipeline(G, 'cuda')
    for i in range(5):
        pipe.submit(ws_host, outs[i & 1])
    pipe.finish()
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)


def test_graphed_synthesis_matches_eager(pkg):
    """networks.GraphedSynthesis: CUDA-graph replay of the synthesis forward returns the eager images, also for new latents."""
    G, g = _build(pkg, 'tinyT')
    ws = cu(g.z['tinyT/ws'])
    gs = pkg.networks.GraphedSynthesis(G.synthesis, ws)
    ref = G.synthesis(ws, noise_mode='const', force_fp32=True)
    assert torch.equal(gs(ws), ref)
    ws2 = ws.flip(0) * 0.9
    ref2 = G.synthesis(ws2, noise_mode='const', force_fp32=True)
    assert torch.equal(gs(ws2), ref2)


@pytest.mark.parametrize('math', ['fp32', 'tf32'])
@pytest.mark.parametrize('name', ['tinyR', 'tinyT'])
def test_pipelined_synthesis_matches_plain_forward(pkg, name, math):
    """networks.PipelinedSynthesis (micro-batches on two streams: convs on a high-priority stream with the shared-memory budget,
    stencils on a normal one): the image equals the plain forward.  Every op is per sample; the only cross-sample term is the
    batch-global style RMS (networks_stylegan3.py:42), which cancels under demodulation up to its 1e-8 epsilon -- so fp32 math agrees
    to rounding, and with TF32 the few weights whose rounding flips stay inside the TF32 network tolerance.  Also: the golden image,
    ragged micro-batches, `out=`, and the budget being restored."""
    from sg3_b200 import capi, modulated_conv, networks
    G, g = _build(pkg, name)
    ws1 = cu(g.z[name + '/ws'])
    ws = torch.cat([ws1, ws1.flip(0) * 0.9, ws1 * 1.1])[:5].contiguous()            # batch 5: ragged micro-batches
    modulated_conv.set_math(math)
    try:
        with torch.no_grad():
            ref = G.synthesis(ws, noise_mode='const', force_fp32=True)
        tol = 1e-5 if math == 'fp32' else 1e-2
        for M in (2, 3, 8):
            pipe = networks.PipelinedSynthesis(G.synthesis, micro_batches=M,
                                               conv_smem_budget=networks.PipelinedSynthesis.CONV_SMEM_BUDGET if M == 2 else 0)
            img = pipe(ws, noise_mode='const', force_fp32=True)
            torch.cuda.synchronize()
            assert img.shape == ref.shape and img.dtype == torch.float32
            assert rel_err(img.cpu().numpy(), ref.cpu().numpy()) < tol, (M, math)
        out = torch.full_like(ref, float('nan'))
        res = networks.PipelinedSynthesis(G.synthesis, micro_batches=2)(ws, out=out, noise_mode='const', force_fp32=True)
        assert res is out and rel_err(out.cpu().numpy(), ref.cpu().numpy()) < tol
        if math == 'fp32':
            n = ws1.shape[0]
            assert rel_err(out[:n].cpu().numpy(), g.z[name + '/img']) < 1e-4        # the golden image of the reference
        assert capi.lib().sg3_modconv_set_smem_budget(0) == 0                         # restored after every call
    finally:
        modulated_conv.set_math(None)


def test_conv_smem_budget_keeps_results(pkg):
    """sg3_modconv_set_smem_budget only changes the depth of the operand ring / the tile plan: same numbers for 1x1 and 3x3 convs."""
    from sg3_b200 import capi, modulated_conv
    rng = np.random.RandomState(3)
    for (I, O, H, k, pad) in ((96, 80, 36, 1, 0), (200, 300, 20, 1, 0), (40, 48, 36, 3, 2), (64, 130, 20, 3, 2)):
        x = cu(rng.randn(2, I, H, H).astype(np.float32))
        w = cu(rng.randn(O, I, k, k).astype(np.float32))
        s = cu(rng.randn(2, I).astype(np.float32))
        modulated_conv.set_math('tf32')
        try:
            y0 = modulated_conv.modulated_conv2d(x, w, s, padding=pad)
            prev = capi.lib().sg3_modconv_set_smem_budget(100 * 1024)
            assert prev == 0
            y1 = modulated_conv.modulated_conv2d(x, w, s, padding=pad)
            assert capi.lib().sg3_modconv_set_smem_budget(0) == 100 * 1024
        finally:
            modulated_conv.set_math(None)
            capi.lib().sg3_modconv_set_smem_budget(0)
        assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-6, (I, O, H, k)


@pytest.mark.parametrize('shape', TC_CASES, ids=['x'.join(map(str, s)) for s in TC_CASES])
def test_modconv_tc_f16_vs_oracle(pkg, shape):
    """fp16 tensor-core contraction (fp16 activations and weights, fp32 accumulate, fp16 store) vs the oracle on the same
    fp16-rounded operands; tolerance 2e-3 of max |ref| (fp16 output rounding)."""
    from oracle import sg3_oracle as orc
    from sg3_b200 import capi
    N, I, O, H, W = shape
    if (H * W) % 8:
        pytest.skip('fp16 TMA needs 16-byte plane pitch')
    rng = np.random.RandomState(I + O)
    x = rng.randn(N, I, H, W).astype(np.float16)
    wm = (rng.randn(N, O, I) / np.sqrt(I)).astype(np.float16)
    ldw = (I + 63) // 64 * 64
    wpad = np.zeros((N, O, ldw), np.float16)
    wpad[:, :, :I] = wm
    xt, wt = cu(x), cu(wpad)
    y = torch.empty(N, O, H, W, device='cuda', dtype=torch.float16)
    rc = capi.lib().sg3_modconv_fwd(xt.data_ptr(), wt.data_ptr(), y.data_ptr(), N, I, O, H, W, 1, 0, ldw, 1, capi.SG3_F16,
                                    capi.stream_ptr(xt.device))
    assert rc == 0
    torch.cuda.synchronize()
    ref = orc.conv2d(x.astype(np.float32), wm.astype(np.float32).reshape(N, O, I, 1, 1), padding=0)
    assert rel_err(y.float().cpu().numpy(), ref) < 2e-3


def test_modconv_half_weights(pkg):
    """The fp16 prologue output equals the fp32 prologue output rounded to fp16."""
    rng = np.random.RandomState(6)
    w, s = rng.randn(37, 70, 1, 1).astype(np.float32), rng.randn(3, 70).astype(np.float32)
    a = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda())[:, :, :70]
    b = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda(), half=True)
    assert b.dtype == torch.float16 and b.shape[2] == 128
    assert torch.equal(a.half(), b[:, :, :70]) and float(b[:, :, 70:].abs().max()) == 0.0


@pytest.mark.parametrize('seed', list(range(12)))
def test_modconv_tc_random_shapes(pkg, seed):
    """Randomised sweep over the tensor-core contractions (1x1 TF32 and 3x3 TF32) through the public modulated_conv2d op:
    ragged channel counts, odd heights, several tile shapes; vs the oracle at the TF32 tolerance."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(500 + seed)
    k = 1 if seed % 2 == 0 else 3
    N, I, O = int(rng.randint(1, 4)), int(rng.randint(3, 200)), int(rng.randint(3, 300))
    H, W = int(rng.randint(5, 90)), 4 * int(rng.randint(2, 60))
    if k == 1 and (H * W) % 4:
        H += 4 - (H % 4)
    x = rng.randn(N, I, H, W).astype(np.float32)
    w = rng.randn(O, I, k, k).astype(np.float32)
    s = rng.randn(N, I).astype(np.float32)
    ref = orc.modulated_conv2d(x, w, s, demodulate=True, padding=k - 1, input_gain=np.float32(0.8))
    before = pkg.capi.lib().sg3_launch_count()
    y = pkg.modulated_conv.modulated_conv2d(cu(x), cu(w), cu(s), demodulate=True, padding=k - 1, input_gain=torch.tensor(0.8).cuda(), math='tf32')
    assert pkg.capi.lib().sg3_modconv_tc_supported(I, O, H, W, k, k - 1) == 0
    assert pkg.capi.lib().sg3_launch_count() - before == 3           # style norm, weight prologue, tensor-core contraction
    assert rel_err(y.cpu().numpy(), ref) < 3e-3


@pytest.mark.parametrize('demod', [True, False])
def test_modconv_weights_backward_kernel(pkg, demod):
    """sg3_modconv_weights_bwd (fused chain rule through the weight prologue) vs torch autograd through the same formula."""
    rng = np.random.RandomState(11)
    N, O, I = 3, 45, 70
    w = cu(rng.randn(O, I, 1, 1).astype(np.float32), True)
    s = cu(rng.randn(N, I).astype(np.float32), True)
    for gain in (None, torch.tensor(0.7).cuda(), cu(rng.rand(I).astype(np.float32) + 0.5)):
        ldw = (I + 31) // 32 * 32
        dWn = torch.zeros(N, O, ldw, device='cuda')
        dWn[:, :, :I] = cu(rng.randn(N, O, I).astype(np.float32))
        # reference: autograd through the prologue expression of networks_stylegan3.py:39-56
        ww, ss = w, s
        if demod:
            ww = ww * ww.square().mean([1, 2, 3], keepdim=True).rsqrt()
            ss = ss * ss.square().mean().rsqrt()
        W = ww.reshape(1, O, I) * ss.unsqueeze(1)
        if demod:
            W = W * (W.square().sum(dim=2, keepdim=True) + 1e-8).rsqrt()
        if gain is not None:
            W = W * gain.expand(N, I).unsqueeze(1)
        dw_ref, ds_ref = torch.autograd.grad(W, [w, s], dWn[:, :, :I])
        dw, ds = pkg.modulated_conv._weights_backward(dWn, w, s, gain, demod, ldw)
        assert rel_err(dw.cpu().numpy(), dw_ref.reshape(O, I).cpu().numpy()) < 2e-5
        assert rel_err(ds.cpu().numpy(), ds_ref.cpu().numpy()) < 2e-5


@pytest.mark.parametrize('case', golden('modconv.npz').cases('modconv'))
def test_modulated_conv2d_fp32x3_golden(pkg, case):
    """3xTF32 contraction (math='fp32x3'): tensor cores with both operands split into TF32 head + tail -- must hold the fp32
    parity tolerance (2e-5), not the TF32 one.  3x3 cases run the exact SIMT kernel under this mode."""
    c = golden('modconv.npz').case('modconv', case)
    k = c['w'].shape[-1]
    g = cu(c['input_gain']) if 'input_gain' in c else None
    y = pkg.modulated_conv.modulated_conv2d(cu(c['x']), cu(c['w']), cu(c['s']), demodulate=bool(c['demodulate']), padding=k - 1,
                                            input_gain=g, math='fp32x3')
    assert rel_err(y.cpu().numpy(), c['y']) < 2e-5


def test_fp32x3_layer_shapes_vs_oracle(pkg):
    """The operand-split kernel at awkward shapes (ragged I / O / P, several N tiles) against the oracle, and that it really is
    the tensor-core launch (launch counter of the library: weights prologue 2 + contraction 1)."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(17)
    for (N, I, O, H) in ((2, 96, 300, 20), (1, 161, 102, 36), (3, 33, 17, 12), (1, 1024, 645, 16)):
        x = rng.randn(N, I, H, H).astype(np.float32) * 3
        w = rng.randn(O, I, 1, 1).astype(np.float32)
        s = (rng.randn(N, I) + 1).astype(np.float32)
        ref = orc.modulated_conv2d(x, w, s)
        n0 = pkg.capi.lib().sg3_launch_count()
        y = pkg.modulated_conv.modulated_conv2d(cu(x), cu(w), cu(s), math='fp32x3')
        assert pkg.capi.lib().sg3_launch_count() - n0 == 3
        assert rel_err(y.cpu().numpy(), ref) < 1e-5, (N, I, O, H)           # fp32 parity tolerance of this suite: 2e-5
        ytf = pkg.modulated_conv.modulated_conv2d(cu(x), cu(w), cu(s), math='tf32')
        assert rel_err(ytf.cpu().numpy(), ref) > rel_err(y.cpu().numpy(), ref)


def test_tiny_generator_fp32x3(pkg):
    """Whole tiny R generator with the 3xTF32 convs: the fp32 network tolerance (1e-4), where TF32 needs 1e-2."""
    G, g = _build(pkg, 'tinyR')
    ws = cu(g.z['tinyR/ws'])
    pkg.modulated_conv.set_math('fp32x3')
    try:
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
    finally:
        pkg.modulated_conv.set_math(None)
    assert rel_err(img.cpu().numpy(), g.z['tinyR/img']) < 1e-4


def test_native_3x3_input_gradient(pkg):
    """dx of the 3x3 modulated conv on the tcgen05 kernel (flipped / transposed taps, padding 2 - pad, row-pitched dy) vs the
    oracle's input gradient; both forward paddings; dy handed over contiguous (re-pitched by a copy) and as a row-pitched view."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(23)
    for (N, I, O, H, pad) in ((2, 40, 24, 24, 2), (1, 96, 130, 20, 2), (2, 33, 64, 20, 0), (1, 32, 40, 24, 2)):   # W % 4 == 0 as in every config-T layer; I = 32: the dgrad conv stacks kx
        x = rng.randn(N, I, H, H).astype(np.float32)
        w = rng.randn(O, I, 3, 3).astype(np.float32)
        s = (rng.randn(N, I) + 1).astype(np.float32)
        OH = H + 2 * pad - 2
        dy = rng.randn(N, O, OH, OH).astype(np.float32)
        dx_ref, dw_ref, ds_ref = orc.modulated_conv2d_bwd(x, w, s, dy, padding=pad)
        for pitched in (False, True):
            xt, wt, st = cu(x, True), cu(w, True), cu(s, True)
            y = pkg.modulated_conv.modulated_conv2d(xt, wt, st, padding=pad, math='tf32')
            if pitched:
                g = pkg.modulated_conv.empty_row_pitched(dy.shape, torch.float32, 'cuda')
                g.copy_(cu(dy))
            else:
                g = cu(dy)
            n0 = pkg.capi.lib().sg3_launch_count()
            dx, dw, ds = torch.autograd.grad(y, [xt, wt, st], g)
            # weight prologue (2) + the tensor-core dgrad conv (1) + the tensor-core weight gradient (1) + the chain rule through the
            # weight prologue (3: style norm, weights, styles): no library convolution, no autograd through the weight expression
            assert pkg.capi.lib().sg3_launch_count() - n0 == 7
            assert rel_err(dx.cpu().numpy(), dx_ref) < 2e-3, (N, I, O, H, pad, pitched)
            assert rel_err(dw.cpu().numpy(), dw_ref) < 2e-3 and rel_err(ds.cpu().numpy(), ds_ref) < 2e-3


@pytest.mark.parametrize('shape', [
    # N, I, O, H, W, pad
    (2, 40, 24, 24, 24, 2),          # OW = 26: dy goes through a row-pitched copy
    (1, 96, 130, 20, 36, 2),         # two o tiles, two i tiles of 48
    (2, 33, 64, 20, 20, 0),          # padding 0 (dgrad-style geometry), one ragged i tile
    (1, 51, 32, 150, 64, 2),         # tall: the row range is split across CTAs (fp32 atomics)
    (1, 16, 16, 9, 132, 2),          # five column chunks, the last one ragged
    (3, 7, 5, 12, 38, 2),            # W % 4 != 0: x is re-pitched too; tiny channel counts
    (1, 200, 51, 16, 16, 2),         # five i tiles
])
def test_wgrad3_vs_oracle(pkg, shape):
    """3x3 weight gradient on the tcgen05 kernel (sg3_modconv_wgrad3: TMA dy tiles, SIMT-shifted x copies, nine TMEM accumulators)
    vs the CPU oracle's conv2d_wgrad; TF32 operands, fp32 accumulation: 2e-3 of max |dw| (stated separately from fp32 parity)."""
    from oracle import sg3_oracle as orc
    N, I, O, H, W, pad = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(N, I, H, W).astype(np.float32)
    dy = rng.randn(N, O, H + 2 * pad - 2, W + 2 * pad - 2).astype(np.float32)
    ref = orc.conv2d_wgrad(x, dy, 3, padding=pad)
    n0 = pkg.capi.lib().sg3_launch_count()
    got = pkg.modulated_conv.conv3x3_weight_grad(cu(x), cu(dy), pad)
    assert got is not None and pkg.capi.lib().sg3_launch_count() - n0 == 1
    assert tuple(got.shape) == ref.shape
    assert rel_err(got.cpu().numpy(), ref) < 2e-3, shape
    # structured input: a single bright pixel of x and of dy picks exactly one tap (catches a shifted or transposed tap index)
    x1 = np.zeros_like(x); dy1 = np.zeros_like(dy)
    x1[0, I - 1, H // 2, W // 2] = 1.0
    for ky in range(3):
        for kx in range(3):
            dy1[:] = 0
            oy, ox = H // 2 - ky + pad, W // 2 - kx + pad
            dy1[0, O - 1, oy, ox] = 1.0
            g1 = pkg.modulated_conv.conv3x3_weight_grad(cu(x1), cu(dy1), pad).cpu().numpy()
            want = np.zeros_like(ref); want[0, O - 1, I - 1, ky, kx] = 1.0
            assert np.array_equal(g1, want), (shape, ky, kx)


def test_wgrad3_rejects_unaddressable(pkg):
    """The C entry point reports NOKERNEL (not a wrong answer) for pitches / pointers TMA cannot address."""
    L = pkg.capi.lib()
    x = torch.zeros(1, 8, 10, 12, device='cuda'); dy = torch.zeros(1, 8, 12, 16, device='cuda'); dw = torch.zeros(1, 9, 8, 8, device='cuda')
    st = pkg.capi.stream_ptr(x.device)
    assert L.sg3_modconv_wgrad3(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), 1, 8, 8, 10, 12, 2, 8, 0, 0, st) == pkg.capi.SG3_E_NOKERNEL   # OW = 14
    assert L.sg3_modconv_wgrad3(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), 1, 8, 8, 10, 12, 2, 8, 16, 0, st) == 0
    assert L.sg3_modconv_wgrad3(dy.data_ptr(), x.data_ptr() + 4, dw.data_ptr(), 1, 8, 8, 10, 12, 2, 8, 16, 0, st) == pkg.capi.SG3_E_NOKERNEL
    assert L.sg3_modconv_wgrad3(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), 1, 8, 8, 10, 12, 1, 8, 16, 0, st) == pkg.capi.SG3_E_NOKERNEL   # pad 1
    torch.cuda.synchronize()


@pytest.mark.parametrize('I,O,P,ldw', [(33, 20, (12, 16), 33), (33, 20, (12, 16), 36), (161, 102, (24, 40), 192), (70, 130, (9, 12), 71)])
def test_wgrad1_c_abi_vs_oracle(pkg, I, O, P, ldw):
    """1x1 weight gradient through the C ABI (sg3_modconv_wgrad): rows of dw that are 16-byte aligned take 128-bit reductions
    (red.global.add.v4.f32), any other ldw the scalar atomics; both against the oracle, padding columns [I, ldw) stay zero."""
    from oracle import sg3_oracle as orc
    rng = np.random.RandomState(I + O + ldw)
    N, (H, W) = 2, P
    x = rng.randn(N, I, H, W).astype(np.float32)
    dy = rng.randn(N, O, H, W).astype(np.float32)
    ref = orc.conv2d_wgrad(x, dy, 1, padding=0).reshape(N, O, I)
    xt, dyt = cu(x), cu(dy)
    dw = torch.zeros(N, O, ldw, device='cuda')
    rc = pkg.capi.lib().sg3_modconv_wgrad(dyt.data_ptr(), xt.data_ptr(), dw.data_ptr(), N, I, O, H, W, ldw, pkg.capi.stream_ptr(xt.device))
    assert rc == 0
    assert rel_err(dw[:, :, :I].cpu().numpy(), ref) < 2e-3
    assert float(dw[:, :, I:].abs().max()) == 0.0 if ldw > I else True


@pytest.mark.parametrize('shape', [
    # N, I, O, H, W, pad
    (2, 40, 24, 24, 24, 2),          # one chunk of 64 channels (40 real), M = 64 path, OW = 26
    (1, 96, 130, 20, 36, 2),         # two o tiles, two chunks, W % 8 == 4 (row-pitched copy of x) like every config-T layer
    (2, 33, 64, 20, 24, 0),          # padding 0
    (1, 200, 51, 30, 148, 2),        # four chunks, three column tiles of 56 / 120 / 184 kept columns
    (1, 64, 32, 70, 276, 2),         # wide rows
])
def test_modconv_tc3_f16_vs_oracle(pkg, shape):
    """fp16 3x3 tensor-core contraction (kind::f16, fp32 accumulation, fp16 store) through the public op vs the oracle on the same
    fp16-rounded activations; the weights go through the fp16 prologue (what `w.to(x.dtype)` does in the reference, :61).
    Tolerance 2e-3 of max |ref| (fp16 weight and output rounding)."""
    from oracle import sg3_oracle as orc
    N, I, O, H, W, pad = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(N, I, H, W).astype(np.float16)
    w = rng.randn(O, I, 3, 3).astype(np.float32)
    s = (rng.randn(N, I) + 1).astype(np.float32)
    ref = orc.modulated_conv2d(x.astype(np.float32), w, s, demodulate=True, padding=pad)
    n0 = pkg.capi.lib().sg3_launch_count()
    y = pkg.modulated_conv.modulated_conv2d(cu(x), cu(w), cu(s), demodulate=True, padding=pad, math='tf32')
    assert pkg.capi.lib().sg3_launch_count() - n0 == 3            # weight prologue (2) + ONE fp16 tensor-core conv, no casts
    assert y.dtype == torch.float16 and tuple(y.shape) == ref.shape
    assert rel_err(y.float().cpu().numpy(), ref) < 2e-3, shape


def test_modconv_half_tap_major_weights(pkg):
    """fp16 tap-major prologue output = the fp32 tap-major output rounded to fp16."""
    rng = np.random.RandomState(8)
    w, s = rng.randn(37, 70, 3, 3).astype(np.float32), rng.randn(3, 70).astype(np.float32)
    a = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda(), tap_major=True)
    b = pkg.modulated_conv.modconv_weights(cu(w), cu(s), True, torch.tensor(0.7).cuda(), tap_major=True, half=True)
    assert b.dtype == torch.float16 and tuple(b.shape) == (3, 9, 37, 128)
    assert torch.equal(a[..., :70].half(), b[..., :70])
