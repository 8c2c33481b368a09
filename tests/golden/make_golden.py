"""Generate tests/golden/*.npz by running the REAL reference (impl='ref', CPU, fp32).

Run in the authoring container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py            # all files
    python tests/golden/make_golden.py ops        # one group: ops | modconv | tiny | r256 | transform

The reference has no tests or fixtures of its own, so these files,
made by importing /root/reference unmodified, are what pins oracle/ and the CUDA path.
Nothing here is imported at test time; tests only read the .npz files.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get('SG3_REF_ROOT', '/root/reference')
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from torch_utils.ops import bias_act, filtered_lrelu, upfirdn2d  # noqa: E402  (reference modules)
from models.stylegan3 import networks_stylegan3 as ref_net  # noqa: E402

torch.set_grad_enabled(True)


def T(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.requires_grad_(grad)


def firwin(numtaps, cutoff, width, fs, radial=False):
    f = ref_net.SynthesisLayer.design_lowpass_filter(numtaps, cutoff, width, fs, radial=radial)
    return None if f is None else f.numpy()


# Filters of the kinds the 1024^2 configs use (values as designed for a mid-network layer).
F_UP2 = firwin(12, 11.3, 2 * 13.0, 128)                      # separable 12 taps
F_UP4 = firwin(24, 11.3, 2 * 13.0, 256)                      # separable 24 taps
F_DN2 = firwin(12, 16.0, 2 * 18.0, 128)                      # separable 12 taps
F_DN4 = firwin(24, 16.0, 2 * 18.0, 256)                      # separable 24 taps
F_DN2R = firwin(12, 16.0, 2 * 18.0, 128, radial=True)        # radial 12x12


def gen_upfirdn(rng, out):
    cases = []
    f_small2d = rng.randn(3, 4).astype(np.float32)
    f_sep5 = rng.randn(5).astype(np.float32)
    specs = [
        # (name, f, up, down, padding, flip, gain, shape)
        ('id', None, 1, 1, 0, False, 1.0, (1, 2, 5, 7)),
        ('sep_up2', F_UP2, 2, 1, [11, 10, 11, 10], False, 4.0, (2, 3, 9, 11)),
        ('sep_up4_crop', F_UP4, 4, 1, [-2, -5, -2, -5], False, 16.0, (1, 2, 12, 10)),
        ('sep_dn2', F_DN2, 1, 2, 0, False, 1.0, (1, 3, 31, 36)),
        ('sep_dn4_flip', F_DN4, 1, 4, [3, 1, 0, 2], True, 1.0, (1, 2, 40, 45)),
        ('full_dn2_radial', F_DN2R, 1, 2, 0, False, 1.0, (2, 2, 30, 33)),
        ('full_3x4_updn', f_small2d, [2, 3], [3, 2], [2, 1, 0, 3], False, 1.5, (1, 2, 8, 9)),
        ('full_3x4_flip', f_small2d, 2, 1, [1, 2, 3, 0], True, 0.5, (1, 1, 6, 5)),
        ('sep5_updn', f_sep5, 3, 2, [4, 4, 2, 3], False, 2.0, (1, 2, 7, 8)),
        ('sep5_negpad', f_sep5, 1, 1, [-1, 3, 2, -2], True, 1.0, (1, 1, 12, 13)),
    ]
    for name, f, up, down, pad, flip, gain, shape in specs:
        x = rng.randn(*shape).astype(np.float32)
        xt = T(x, True)
        y = upfirdn2d.upfirdn2d(xt, None if f is None else T(f), up=up, down=down, padding=pad,
                                flip_filter=flip, gain=gain, impl='ref')
        dy = rng.randn(*y.shape).astype(np.float32)
        (dx,) = torch.autograd.grad(y, xt, T(dy))
        k = f'upfirdn/{name}/'
        out[k + 'x'] = x
        if f is not None:
            out[k + 'f'] = f
        out[k + 'up'] = np.asarray(up if isinstance(up, list) else [up, up])
        out[k + 'down'] = np.asarray(down if isinstance(down, list) else [down, down])
        out[k + 'padding'] = np.asarray(pad if isinstance(pad, list) else [pad] * 4)
        out[k + 'flip'] = np.asarray(flip)
        out[k + 'gain'] = np.asarray(gain)
        out[k + 'y'] = y.detach().numpy()
        out[k + 'dy'] = dy
        out[k + 'dx'] = dx.numpy()
        cases.append(name)
    out['upfirdn/_cases'] = np.asarray(cases)


def gen_bias_act(rng, out):
    cases = []
    for act in bias_act.activation_funcs.keys():
        for variant, (shape, dim, clamp, gain, alpha) in {
            'a': ((3, 5, 4, 6), 1, None, None, None),
            'b': ((4, 7), 1, 0.9, 1.7, 0.3),
            'c': ((2, 3, 5), 2, 0.5, None, None),
        }.items():
            x = (rng.randn(*shape) * 1.5).astype(np.float32)
            b = rng.randn(shape[dim]).astype(np.float32)
            dy = rng.randn(*shape).astype(np.float32)
            ddx = rng.randn(*shape).astype(np.float32)
            xt, bt, dyt = T(x, True), T(b, True), T(dy, True)
            y = bias_act.bias_act(xt, bt, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref')
            dx, db = torch.autograd.grad(y, [xt, bt], dyt, create_graph=True)
            g2 = torch.autograd.grad(dx, [dyt, xt], T(ddx), allow_unused=True)
            k = f'bias_act/{act}_{variant}/'
            out[k + 'x'], out[k + 'b'], out[k + 'dy'], out[k + 'ddx'] = x, b, dy, ddx
            out[k + 'dim'] = np.asarray(dim)
            out[k + 'clamp'] = np.asarray(-1.0 if clamp is None else clamp)
            out[k + 'gain'] = np.asarray(np.nan if gain is None else gain)
            out[k + 'alpha'] = np.asarray(np.nan if alpha is None else alpha)
            out[k + 'y'] = y.detach().numpy()
            out[k + 'dx'] = dx.detach().numpy()
            out[k + 'db'] = db.detach().numpy()
            out[k + 'd_dy'] = g2[0].numpy()
            out[k + 'd_x'] = (g2[1] if g2[1] is not None else torch.zeros_like(xt)).numpy()
            cases.append(f'{act}_{variant}')
    out['bias_act/_cases'] = np.asarray(cases)


def gen_flrelu(rng, out):
    cases = []
    f2 = rng.randn(4, 4).astype(np.float32) * 0.3
    specs = [
        # name, fu, fd, up, down, padding, gain, slope, clamp, flip, shape
        ('R_same', F_UP2, F_DN2R, 2, 2, [11, 10, 11, 10], np.sqrt(2), 0.2, 256, False, (2, 3, 20, 20)),
        ('R_up4', F_UP4, F_DN2R, 4, 2, [-2, -5, -2, -5], np.sqrt(2), 0.2, 256, False, (1, 3, 20, 20)),
        ('T_same', F_UP2, F_DN2, 2, 2, [9, 8, 9, 8], np.sqrt(2), 0.2, 256, False, (2, 2, 22, 22)),
        ('T_up4', F_UP4, F_DN2, 4, 2, [-6, -9, -6, -9], np.sqrt(2), 0.2, 256, False, (1, 2, 22, 22)),
        ('crit_last', F_UP2, F_DN2, 2, 2, [-9, -10, -9, -10], np.sqrt(2), 0.2, 256, False, (1, 2, 36, 36)),
        ('torgb', None, None, 1, 1, 0, 1.0, 1.0, 256, False, (2, 3, 16, 16)),
        ('clampy', F_UP2, F_DN2, 2, 2, [11, 10, 11, 10], np.sqrt(2), 0.2, 0.75, False, (1, 2, 17, 19)),
        ('rect_flip', F_UP2, F_DN4, 2, 4, [7, 9, 12, 6], 1.3, 0.1, 2.0, True, (1, 2, 21, 33)),
        ('full_up_sep_dn', F_DN2R, F_UP2, 2, 2, [10, 11, 10, 11], 1.0, 0.2, None, True, (1, 2, 18, 18)),
        ('generic_3_2', f2, f2, 3, 2, [2, 3, 1, 4], 1.1, 0.3, 1.5, False, (1, 2, 9, 10)),
        ('up1_dn2', None, F_DN2, 1, 2, [0, 0, 0, 0], np.sqrt(2), 0.2, 1.0, False, (1, 2, 30, 31)),
        ('up2_dn1', F_UP2, None, 2, 1, [6, 5, 6, 5], np.sqrt(2), 0.2, 1.0, False, (1, 2, 12, 13)),
        # appended in round 2 (the rng stream of the cases above is unchanged): a missing filter next to a factor of 2 -- fu=None is
        # the 1x1 identity (zero insertion only), fd=None plain decimation; the backward of the second is the first
        ('fuNone_up2_dn2', None, F_DN2, 2, 2, [3, 2, 3, 2], np.sqrt(2), 0.2, 1.0, False, (1, 2, 20, 21)),
        ('up2_dn2_fdNone', F_UP2, None, 2, 2, [6, 5, 6, 5], np.sqrt(2), 0.2, 1.0, False, (1, 2, 14, 15)),
    ]
    for name, fu, fd, up, down, pad, gain, slope, clamp, flip, shape in specs:
        x = (rng.randn(*shape) * (4.0 if clamp and clamp < 10 else 2.0)).astype(np.float32)
        if name in ('R_same', 'torgb'):
            x *= 80.0  # make the 256 clamp bite
        b = rng.randn(shape[1]).astype(np.float32)
        xt, bt = T(x, True), T(b, True)
        y = filtered_lrelu.filtered_lrelu(xt, None if fu is None else T(fu), None if fd is None else T(fd), bt,
                                          up=up, down=down, padding=pad, gain=gain, slope=slope, clamp=clamp,
                                          flip_filter=flip, impl='ref')
        dy = rng.randn(*y.shape).astype(np.float32)
        dx, db = torch.autograd.grad(y, [xt, bt], T(dy))
        k = f'flrelu/{name}/'
        out[k + 'x'], out[k + 'b'], out[k + 'dy'] = x, b, dy
        if fu is not None:
            out[k + 'fu'] = fu
        if fd is not None:
            out[k + 'fd'] = fd
        out[k + 'up'], out[k + 'down'] = np.asarray(up), np.asarray(down)
        out[k + 'padding'] = np.asarray(pad if isinstance(pad, list) else [pad] * 4)
        out[k + 'gain'], out[k + 'slope'] = np.asarray(float(gain)), np.asarray(float(slope))
        out[k + 'clamp'] = np.asarray(-1.0 if clamp is None else float(clamp))
        out[k + 'flip'] = np.asarray(flip)
        out[k + 'y'] = y.detach().numpy()
        out[k + 'dx'], out[k + 'db'] = dx.numpy(), db.numpy()
        cases.append(name)
    out['flrelu/_cases'] = np.asarray(cases)


def gen_modconv(rng, out):
    cases = []
    specs = [
        # name, N, I, O, k, H, W, demod, gain_kind
        ('k1_demod', 3, 24, 20, 1, 9, 11, True, 'scalar'),
        ('k3_demod', 2, 10, 12, 3, 8, 7, True, 'scalar'),
        ('k1_torgb', 2, 16, 3, 1, 12, 12, False, 'scalar'),
        ('k3_nogain', 2, 6, 5, 3, 6, 6, True, None),
        ('k1_gain_ni', 2, 8, 9, 1, 5, 6, True, 'ni'),
    ]
    for name, N, I, O, k, H, W, demod, gk in specs:
        x = rng.randn(N, I, H, W).astype(np.float32)
        w = rng.randn(O, I, k, k).astype(np.float32)
        s = (rng.randn(N, I) * 0.5 + 1.0).astype(np.float32)
        g = None
        if gk == 'scalar':
            g = np.asarray(0.8, np.float32)
        elif gk == 'ni':
            g = (rng.rand(N, I) + 0.5).astype(np.float32)
        xt, wt, st = T(x, True), T(w, True), T(s, True)
        y = ref_net.modulated_conv2d(xt, wt, st, demodulate=demod, padding=k - 1,
                                     input_gain=None if g is None else T(g))
        dy = rng.randn(*y.shape).astype(np.float32)
        dx, dw, ds = torch.autograd.grad(y, [xt, wt, st], T(dy))
        kk = f'modconv/{name}/'
        out[kk + 'x'], out[kk + 'w'], out[kk + 's'], out[kk + 'dy'] = x, w, s, dy
        if g is not None:
            out[kk + 'input_gain'] = g
        out[kk + 'demodulate'] = np.asarray(demod)
        out[kk + 'y'] = y.detach().numpy()
        out[kk + 'dx'], out[kk + 'dw'], out[kk + 'ds'] = dx.numpy(), dw.numpy(), ds.numpy()
        cases.append(name)
    out['modconv/_cases'] = np.asarray(cases)


TINY = dict(
    tinyR=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=1, use_radial_filters=True),
    tinyT=dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
               conv_kernel=3, use_radial_filters=False),
)


def gen_tiny(out):
    for name, cfg in TINY.items():
        torch.manual_seed(0)
        G = ref_net.Generator(**cfg).eval().requires_grad_(False)
        z = torch.randn(2, cfg['z_dim'], generator=torch.Generator().manual_seed(1))
        ws = G.mapping(z, None)
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
        # Intermediate activations after a few layers, for layer-level parity.
        feats = {}
        x = G.synthesis.input(ws[:, 0])
        feats['input'] = x[:1, :8].contiguous().numpy()   # sample 0, first 8 channels keep the file small
        for i, (lname, w) in enumerate(zip(G.synthesis.layer_names, ws.unbind(1)[1:])):
            x = getattr(G.synthesis, lname)(x, w, noise_mode='const', force_fp32=True)
            if i in (0, 2, 7, 12):
                feats[f'after_{i}'] = x[:1, :8].contiguous().numpy()
        k = f'{name}/'
        for sk, sv in G.synthesis.state_dict().items():
            out[k + 'state/' + sk] = sv.numpy()
        out[k + 'ws'] = ws.numpy()
        out[k + 'z'] = z.numpy()
        out[k + 'img'] = img.numpy()
        for fk, fv in feats.items():
            out[k + 'feat/' + fk] = fv
        # Gradients through the whole synthesis network (PTI-style loss = mean(img*t)).
        G.requires_grad_(True)
        tgt = torch.randn(img.shape, generator=torch.Generator().manual_seed(2))
        loss = (G.synthesis(ws, noise_mode='const', force_fp32=True) * tgt).mean()
        params = dict(G.synthesis.named_parameters())
        grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
        out[k + 'tgt'] = tgt.numpy()
        out[k + 'loss'] = loss.detach().numpy()
        for (pk, _), g in zip(params.items(), grads):
            if g is not None:
                out[k + 'grad/' + pk] = g.numpy()
        print(name, 'img mean/std/absmax', float(img.mean()), float(img.std()), float(img.abs().max()))


def gen_r256(out):
    """Config 1 (BASELINE.json configs[0]): StyleGAN3-R 256^2, random init seed 0, batch 4 -- ~75 s on 8 cores."""
    cfg = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=256, img_channels=3, channel_base=65536,
               channel_max=1024, conv_kernel=1, use_radial_filters=True)
    torch.manual_seed(0)
    G = ref_net.Generator(**cfg).eval().requires_grad_(False)
    z = torch.randn(4, 512)
    ws = G.mapping(z, None)
    img = G.synthesis(ws, noise_mode='const', force_fp32=True)
    out['r256/ws'] = ws.numpy()
    out['r256/img_sub4'] = img[:, :, ::4, ::4].contiguous().numpy()     # strided subsample keeps the file small
    out['r256/stats'] = np.asarray([float(img.mean()), float(img.std()), float(img.abs().max())])
    # a few weights to prove the package's seed-0 construction consumes the RNG identically
    sd = G.synthesis.state_dict()
    names = G.synthesis.layer_names
    for key in ['input.weight', 'input.freqs', names[0] + '.weight', names[13] + '.weight', names[14] + '.affine.weight']:
        out['r256/probe/' + key] = sd[key].flatten()[:64].numpy()
    print('r256 stats', out['r256/stats'])


def gen_transform(out):
    """Per-sample user transforms on `synthesis.input.transform` ([N, 3, 3]: what the ReStyle / pSp inversion sets from the
    landmarks, models/setgan/encoder/psp3.py:63-65, and the video FOV expansion translates, utils/fov_expansion.py:20-27) and a
    single [3, 3] transform (the reference's default buffer shape).  Weights = the tiny.npz generators (same seed)."""
    for name, cfg in TINY.items():
        torch.manual_seed(0)
        G = ref_net.Generator(**cfg).eval().requires_grad_(False)
        ws = G.mapping(torch.randn(2, cfg['z_dim'], generator=torch.Generator().manual_seed(1)), None)
        rng = np.random.RandomState(7)
        mats = []
        for ang, sc, tx, ty in ((0.3, 1.15, 0.12, -0.07), (-1.1, 0.8, -0.3, 0.25)):
            c, s = np.cos(ang) * sc, np.sin(ang) * sc
            mats.append(np.array([[c, -s, tx], [s, c, ty], [0, 0, 1]], np.float32))
        for tag, m in (('per_sample', np.stack(mats)), ('single', mats[0])):
            G.synthesis.input.transform = torch.from_numpy(m)
            x = G.synthesis.input(ws[:, 0])
            img = G.synthesis(ws, noise_mode='const', force_fp32=True)
            k = f'{name}/{tag}/'
            out[k + 'transform'] = m
            out[k + 'input'] = x[:, :8].contiguous().numpy()
            out[k + 'img'] = img.numpy()
            print(name, tag, 'img absmax', float(img.abs().max()))
        out[f'{name}/ws'] = ws.numpy()


def main():
    which = sys.argv[1:] or ['ops', 'modconv', 'tiny', 'r256', 'transform']
    torch.set_num_threads(os.cpu_count())
    if 'ops' in which:
        out = {}
        rng = np.random.RandomState(1234)
        gen_upfirdn(rng, out)
        gen_bias_act(rng, out)
        gen_flrelu(rng, out)
        np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)
    if 'modconv' in which:
        out = {}
        gen_modconv(np.random.RandomState(99), out)
        np.savez_compressed(os.path.join(HERE, 'modconv.npz'), **out)
    if 'tiny' in which:
        out = {}
        gen_tiny(out)
        np.savez_compressed(os.path.join(HERE, 'tiny.npz'), **out)
    if 'transform' in which:
        out = {}
        gen_transform(out)
        np.savez_compressed(os.path.join(HERE, 'transform.npz'), **out)
    if 'r256' in which:
        out = {}
        gen_r256(out)
        np.savez_compressed(os.path.join(HERE, 'r256.npz'), **out)


if __name__ == '__main__':
    main()
