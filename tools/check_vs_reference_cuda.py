"""Parity and timing against the reference's OWN CUDA plugins on the same B200 (north_star: "match the reference's own
CUDA plugin"; SURVEY 8d: the kernel to beat).

    python tools/check_vs_reference_cuda.py [--n 2] [--iters 10] [--configs R,T] [--no-net] [--conv-batch 32] [--out FILE.md]

Needs a CUDA device and baseline/_ref/ (git-ignored; `sh tools/stage_reference.sh` in the authoring container stages the
UNMODIFIED reference files `torch_utils/custom_ops.py`, `torch_utils/ops/*`, `models/stylegan3/*` there and prebuilds the three
plugins from them with the reference's own flags).  The prebuilt `.so` files are handed to the reference's plugin loader
through its own cache dict (`custom_ops._cached_plugins`), i.e. the reference's Python ops run unmodified on the reference's
CUDA kernels; without a prebuilt module `custom_ops.get_plugin` JIT-builds as usual (~2-3 min).  sg3_b200.install() is NOT
called here: `torch_utils.ops.*` are the reference's modules, `sg3_b200.*` ours, side by side in one process.

Per layer of StyleGAN3-R 1024^2 and StyleGAN3-T 1024^2 (real channel counts, plane sizes, filters, paddings, batch --n):
  parity   y (no grad), y / dx / db with autograd, the sign tensor (2-bit codes that differ), bias_act / upfirdn2d spot checks;
  timing   CUDA events, ours vs the reference plugin: forward, forward with sign write, backward.
Whole generators (reference networks_stylegan3.Generator on the reference plugins + cuDNN vs sg3_b200.networks.Generator, same
seed-0 weights): image parity with fp32 convs, forward time at batch 4 with TF32 convs.
Per-layer modulated_conv2d at --conv-batch: the reference's function (eager weight chain + cuDNN grouped conv) vs ours.
"""
import argparse
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get('SG3_REF_ROOT') or os.path.join(ROOT, 'baseline', '_ref')

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=2)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--configs', default='R,T')
ap.add_argument('--no-net', action='store_true')
ap.add_argument('--conv-batch', type=int, default=32)
ap.add_argument('--out', default=None)
args = ap.parse_args()

if not os.path.exists(os.path.join(REF, 'torch_utils', 'ops', 'filtered_lrelu.py')):
    print(f'reference ops not staged at {REF} (run tools/stage_reference.sh in the authoring container): nothing checked')
    sys.exit(0)
sys.path.insert(0, REF)

import numpy as np
import torch

assert torch.cuda.is_available(), 'needs a CUDA device'
from torch_utils import custom_ops                      # noqa: E402  the reference's loader, unmodified

for name in ('bias_act_plugin', 'upfirdn2d_plugin', 'filtered_lrelu_plugin'):
    so = os.path.join(REF, '_plugins', name, name + '.so')
    if os.path.exists(so):
        spec = importlib.util.spec_from_file_location(name, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[name] = mod
        custom_ops._cached_plugins[name] = mod
        print(f'reference plugin {name}: prebuilt {os.path.relpath(so, ROOT)}')
    else:
        print(f'reference plugin {name}: no prebuilt module, custom_ops.get_plugin will JIT-build it')

from torch_utils.ops import bias_act as ref_ba, filtered_lrelu as ref_fl, upfirdn2d as ref_up      # noqa: E402
import sg3_b200                                                                                      # noqa: E402
from sg3_b200 import bias_act as our_ba, filtered_lrelu as our_fl, modulated_conv, networks, upfirdn2d as our_up   # noqa: E402

assert ref_fl is not our_fl and ref_fl.__file__.startswith(REF)
our_fl._quiet_fallback = True
dev = torch.device('cuda')
lines = []


def out(s=''):
    print(s, flush=True)
    lines.append(s)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


CFG = dict(
    R=dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=65536, channel_max=1024,
           conv_kernel=1, use_radial_filters=True),
    T=dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512,
           conv_kernel=3, use_radial_filters=False),
)

out('# Ours vs the reference\'s own CUDA plugins on the same B200')
out()
out(f'`{torch.cuda.get_device_name()}`, torch {torch.__version__}; reference plugins compiled from the unmodified sources '
    f'(`--use_fast_math`, sm_100); `{sg3_b200.capi.lib().sg3_build_info().decode()}`')
out()
assert ref_fl._init() and ref_ba._init() and ref_up._init()

worst = dict(y=0.0, dx=0.0, db=0.0)
total_ms = {}
for cname in args.configs.split(','):
    torch.manual_seed(0)
    G = networks.Generator(**CFG[cname])
    out(f'## filtered_lrelu, StyleGAN3-{cname} 1024^2 layer shapes, batch {args.n}, fp32')
    out()
    out('| layer | C | in -> out | up,down | y err | y(grad) err | dx err | dx err, same signs | db err | sign codes differing | fwd ours / ref ms | fwd+signs ours / ref ms | bwd ours / ref ms |')
    out('|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    tot = np.zeros(6)
    for li, lname in enumerate(G.synthesis.layer_names):
        L = getattr(G.synthesis, lname)
        C = L.out_channels
        size = int(L.in_size[0]) + L.conv_kernel - 1
        fu = None if L.up_filter is None else L.up_filter.to(dev)
        fd = None if L.down_filter is None else L.down_filter.to(dev)
        kw = dict(up=L.up_factor, down=L.down_factor, padding=L.padding, gain=(1 if L.is_torgb else np.sqrt(2)),
                  slope=(1 if L.is_torgb else 0.2), clamp=L.conv_clamp)
        g = torch.Generator(device=dev).manual_seed(100 + li)
        x = torch.randn(args.n, C, size, size, device=dev, generator=g) * 4
        x.view(-1)[::9973] *= 80            # a sprinkling of values beyond +-clamp / gain
        b = torch.randn(C, device=dev, generator=g)
        with torch.no_grad():
            y_ref = ref_fl.filtered_lrelu(x, fu, fd, b, **kw)
            y_our = our_fl.filtered_lrelu(x, fu, fd, b, **kw)
        e_y = rel(y_our, y_ref)
        dy = torch.randn(y_ref.shape, device=dev, generator=g)
        res = {}
        for tag, mod in (('ref', ref_fl), ('our', our_fl)):
            xg = x.clone().requires_grad_(True)
            bg = b.clone().requires_grad_(True)
            y = mod.filtered_lrelu(xg, fu, fd, bg, **kw)
            dx, db = torch.autograd.grad(y, [xg, bg], dy)
            res[tag] = (y.detach(), dx, db)
            del xg, bg, y
        e_yg = rel(res['our'][0], res['ref'][0])
        e_dx = rel(res['our'][1], res['ref'][1])
        e_db = rel(res['our'][2], res['ref'][2])
        # sign tensors: the plugin entry point itself vs our fused launch (same layout: uint8 [N, C, sH, ceil16(sW)/4])
        px0, px1, py0, py1 = ref_fl._parse_padding(L.padding)
        e_dxs = e_dx
        if L.is_torgb or fu is None:
            ndiff, ncodes = 0, 0
        else:
            _, so_ref, rc = ref_fl._plugin.filtered_lrelu(x, fu, fd, b, torch.empty([0]), L.up_factor,
                                                          L.down_factor, px0, px1, py0, py1, 0, 0, float(kw['gain']), float(kw['slope']),
                                                          float(kw['clamp']), False, True)
            assert rc == 0, f'reference plugin has no kernel for {lname}'
            cfg = (L.up_factor, L.down_factor, px0, px1, py0, py1, float(kw['gain']), float(kw['slope']), float(kw['clamp']), False)
            _, so_our = our_fl._fused(x, fu, fd, b, None, 0, 0, cfg, True)
            assert so_our.shape == so_ref.shape, (so_our.shape, so_ref.shape)
            # compare only the active region: the padding pixels of the last byte group are unspecified in both
            yh, yw = y_ref.shape[2], y_ref.shape[3]
            fdw = fd.shape[-1]
            sw = yw * L.down_factor - (L.down_factor - 1) + fdw - 1
            shifts = torch.tensor([0, 2, 4, 6], device=dev, dtype=torch.uint8)
            codes_r = ((so_ref.unsqueeze(-1) >> shifts) & 3).reshape(*so_ref.shape[:3], -1)[..., :sw]
            codes_o = ((so_our.unsqueeze(-1) >> shifts) & 3).reshape(*so_our.shape[:3], -1)[..., :sw]
            ndiff, ncodes = int((codes_r != codes_o).sum()), codes_r.numel()
            # the reference plugin's backward launch (roles swapped, filtered_lrelu.py:254-264) reading OUR sign tensor: what is left
            # of the dx difference once the handful of threshold codes is taken out
            fuw, fdw2 = fu.shape[-1], fd.shape[-1]
            xw, yw2 = x.shape[3], y_ref.shape[3]
            pp = [(fuw - 1) + (fdw2 - 1) - px0, xw * L.up_factor - yw2 * L.down_factor + px0 - (L.up_factor - 1),
                  (fuw - 1) + (fdw2 - 1) - py0, xw * L.up_factor - yw2 * L.down_factor + py0 - (L.up_factor - 1)]
            gg = float(kw['gain']) * (L.up_factor ** 2) / (L.down_factor ** 2)
            dx_rs, _, rc = ref_fl._plugin.filtered_lrelu(dy, fd, fu, torch.zeros([C], device=dev), so_our, L.down_factor, L.up_factor,
                                                         pp[0], pp[1], pp[2], pp[3], -(fuw - 1) + px0, -(fuw - 1) + py0, gg,
                                                         float(kw['slope']), float('inf'), True, False)
            assert rc == 0
            e_dxs = rel(res['our'][1], dx_rs)
            del so_ref, so_our, codes_r, codes_o, dx_rs
        for k, v in (('y', max(e_y, e_yg)), ('dx', e_dxs), ('db', e_db)):
            worst[k] = max(worst[k], v)
        # timing
        with torch.no_grad():
            t_of = timeit(lambda: our_fl.filtered_lrelu(x, fu, fd, b, **kw), args.iters)
            t_rf = timeit(lambda: ref_fl.filtered_lrelu(x, fu, fd, b, **kw), args.iters)
        xg = x.clone().requires_grad_(True)
        t_ow = timeit(lambda: our_fl.filtered_lrelu(xg, fu, fd, b, **kw), args.iters)
        t_rw = timeit(lambda: ref_fl.filtered_lrelu(xg, fu, fd, b, **kw), args.iters)
        bg = b.clone().requires_grad_(True)
        yo = our_fl.filtered_lrelu(xg, fu, fd, bg, **kw)
        t_ob = timeit(lambda: torch.autograd.grad(yo, [xg, bg], dy, retain_graph=True), args.iters)
        yr = ref_fl.filtered_lrelu(xg, fu, fd, bg, **kw)
        t_rb = timeit(lambda: torch.autograd.grad(yr, [xg, bg], dy, retain_graph=True), args.iters)
        tot += np.array([t_of, t_rf, t_ow, t_rw, t_ob, t_rb])
        out(f'| {lname} | {C} | {size} -> {y_ref.shape[-1]} | {L.up_factor},{L.down_factor} | {e_y:.1e} | {e_yg:.1e} | {e_dx:.1e} | {e_dxs:.1e} | {e_db:.1e} | '
            f'{ndiff} / {ncodes} | {t_of:.3f} / {t_rf:.3f} | {t_ow:.3f} / {t_rw:.3f} | {t_ob:.3f} / {t_rb:.3f} |')
        del x, y_ref, y_our, dy, res, xg, bg, yo, yr
        torch.cuda.empty_cache()
    out(f'| **sum** | | | | | | | | | | **{tot[0]:.2f} / {tot[1]:.2f}** ({tot[1] / tot[0]:.2f}x) | **{tot[2]:.2f} / {tot[3]:.2f}** ({tot[3] / tot[2]:.2f}x) | '
        f'**{tot[4]:.2f} / {tot[5]:.2f}** ({tot[5] / tot[4]:.2f}x) |')
    out()
    total_ms[cname] = tot
    del G

out(f'Worst relative error (max |ours - ref| / max |ref|) over all layers: y {worst["y"]:.1e}, dx on the same sign tensor {worst["dx"]:.1e}, '
    f'db {worst["db"]:.1e}.  Column "dx err" compares the two autograd paths end to end: each side reads its own sign tensor, and the '
    'handful of 2-bit codes that differ (activations within fp32 rounding of 0 or +-clamp; the reference plugin is built with '
    '`--use_fast_math`) each move one upsampled pixel of the gradient by up to (1 - slope) of its value -- a few percent of max |dx| '
    'at that pixel, nothing elsewhere.  "dx err, same signs" is the reference plugin\'s own backward launch reading our sign tensor.')
out()

# ---- bias_act / upfirdn2d spot checks against the plugins ----
out('## bias_act, upfirdn2d vs the reference plugins')
out()
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(8, 512, device=dev, generator=g)
b = torch.randn(512, device=dev, generator=g)
errs = []
for act in ref_ba.activation_funcs:
    errs.append(rel(our_ba.bias_act(x, b, act=act), ref_ba.bias_act(x, b, act=act)))
out(f'* bias_act [8, 512], all {len(errs)} activations: max err {max(errs):.1e}')
x = torch.randn(2, 16, 300, 300, device=dev, generator=g)
for f, up, down, pad in ((ref_up.setup_filter([1, 3, 3, 1]), 2, 1, [2, 1, 2, 1]), (ref_up.setup_filter([1, 3, 3, 1]), 1, 2, [1, 1, 1, 1]),
                         (torch.randn(12, device='cpu', generator=torch.Generator().manual_seed(1)), 2, 1, [5, 6, 5, 6]),
                         (torch.randn(12, 12, device='cpu', generator=torch.Generator().manual_seed(2)), 1, 2, [0, 0, 0, 0])):
    f = f.to(dev)
    e = rel(our_up.upfirdn2d(x, f, up=up, down=down, padding=pad), ref_up.upfirdn2d(x, f, up=up, down=down, padding=pad))
    out(f'* upfirdn2d [2,16,300,300] f{list(f.shape)} up {up} down {down}: err {e:.1e}')
    assert e < 1e-4
out()

# ---- whole generators ----
if not args.no_net:
    from models.stylegan3 import networks_stylegan3 as ref_net          # noqa: E402  reference model code on the reference plugins
    assert ref_net.filtered_lrelu is ref_fl
    out('## Whole generator: reference model code + reference plugins + cuDNN vs sg3_b200')
    out()
    out('| config | image err, fp32 convs (batch 1) | image err, TF32 convs | forward ms, batch 4: reference / ours | speed-up |')
    out('|---|---|---|---|---|')
    conv_tables = []
    for cname in args.configs.split(','):
        torch.manual_seed(0)
        Gr = ref_net.Generator(**CFG[cname]).eval().requires_grad_(False).to(dev)
        torch.manual_seed(0)
        Go = networks.Generator(**CFG[cname]).eval().requires_grad_(False).to(dev)
        Go.load_state_dict(Gr.state_dict())
        z = torch.randn(4, 512, generator=torch.Generator().manual_seed(1)).to(dev)
        with torch.no_grad():
            ws = Gr.mapping(z, None)
            torch.backends.cudnn.allow_tf32 = False
            modulated_conv.set_math('fp32')
            e32 = rel(Go.synthesis(ws[:1], noise_mode='const', force_fp32=True), Gr.synthesis(ws[:1], noise_mode='const', force_fp32=True))
            torch.backends.cudnn.allow_tf32 = True
            modulated_conv.set_math('tf32')
            img_r = Gr.synthesis(ws[:1], noise_mode='const', force_fp32=True)
            etf = rel(Go.synthesis(ws[:1], noise_mode='const', force_fp32=True), img_r)
            t_r = timeit(lambda: Gr.synthesis(ws, noise_mode='const', force_fp32=True), 3)
            t_o = timeit(lambda: Go.synthesis(ws, noise_mode='const', force_fp32=True), 3)
        out(f'| StyleGAN3-{cname} 1024^2 | {e32:.1e} | {etf:.1e} | {t_r:.1f} / {t_o:.1f} | {t_r / t_o:.2f}x |')
        # per-layer modulated_conv2d at the bench batch: the reference's function (eager weight chain + cuDNN grouped conv)
        if args.conv_batch > 0:
            B = args.conv_batch
            rows = []
            tr_sum = to_sum = 0.0
            for lname in Go.synthesis.layer_names:
                L = getattr(Go.synthesis, lname)
                size = int(L.in_size[0])
                x = torch.randn(B, L.in_channels, size, size, device=dev)
                s = torch.randn(B, L.in_channels, device=dev) + 1
                kwc = dict(padding=L.conv_kernel - 1, demodulate=not L.is_torgb, input_gain=torch.ones([], device=dev))
                with torch.no_grad():
                    try:
                        t_rc = timeit(lambda: ref_net.modulated_conv2d(x=x, w=L.weight, s=s, **kwc), 3)
                    except torch.OutOfMemoryError:
                        t_rc = float('nan')
                    t_oc = timeit(lambda: modulated_conv.modulated_conv2d(x=x, w=L.weight, s=s, **kwc), 3)
                rows.append(f'| {lname} | {L.in_channels}->{L.out_channels} k{L.conv_kernel} @{size} | {t_rc:.3f} | {t_oc:.3f} | {t_rc / t_oc:.2f}x |')
                tr_sum += t_rc
                to_sum += t_oc
                del x, s
                torch.cuda.empty_cache()
            conv_tables.append((cname, B, rows, tr_sum, to_sum))
        del Gr, Go
        torch.cuda.empty_cache()
    out()
    for cname, B, rows, tr_sum, to_sum in conv_tables:
        out(f'### modulated_conv2d per layer, StyleGAN3-{cname}, batch {B}, TF32 (reference = its eager weight chain + cuDNN grouped conv)')
        out()
        out('| layer | shape | reference ms | ours ms | ratio |')
        out('|---|---|---|---|---|')
        for r in rows:
            out(r)
        out(f'| **sum** | | **{tr_sum:.2f}** | **{to_sum:.2f}** | **{tr_sum / to_sum:.2f}x** |')
        out()

if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, 'w') as f:
        f.write('\n'.join(lines) + '\n')
ok = worst['y'] < 1e-3 and worst['db'] < 1e-3 and worst['dx'] < 1e-3
print('PARITY', 'OK' if ok else 'FAILED', worst)
sys.exit(0 if ok else 1)
