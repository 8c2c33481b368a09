"""CPU-side checks (no GPU, no compute launches): the C-ABI library loads and exports every symbol
that include/sg3_b200.h declares, geometry helpers agree with the oracle, host-side argument logic
and the drop-in seam behave like the reference interface."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope='module')
def pkg():
    import __graft_entry__ as ge
    ge.build()
    import sg3_b200
    return sg3_b200


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, 'include', 'sg3_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = sorted(set(re.findall(r'\b(sg3_[a-z0-9_]+)\s*\(', hdr)))
    assert len(declared) >= 10
    lib = ctypes.CDLL(pkg.capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in sg3_b200.h but not exported'
    assert sorted(pkg.capi.EXPORTS) == declared
    assert pkg.capi.lib().sg3_abi_version() == 1
    assert b'sm_100a' in pkg.capi.lib().sg3_build_info()


def test_struct_layout_matches_header(pkg):
    assert ctypes.sizeof(pkg.capi.FlreluDesc) == pkg.capi.lib().sg3_sizeof_flrelu_desc()
    assert pkg.capi.FlreluDesc.xStride.offset == 6 * 8 + 6 * 4 and pkg.capi.FlreluDesc.dtype.offset == 212


@pytest.mark.parametrize('cfg', [
    dict(inH=36, inW=36, up=2, down=2, fu=(12, 0), fd=(12, 12), pad=[11, 10, 11, 10]),
    dict(inH=84, inW=84, up=4, down=2, fu=(24, 0), fd=(12, 12), pad=[-2, -5, -2, -5]),
    dict(inH=1044, inW=1044, up=2, down=2, fu=(12, 0), fd=(12, 0), pad=[-9, -10, -9, -10]),
    dict(inH=17, inW=23, up=1, down=1, fu=(1, 1), fd=(1, 1), pad=[0, 0, 0, 0]),
    dict(inH=21, inW=33, up=2, down=4, fu=(12, 0), fd=(24, 0), pad=[7, 9, 12, 6]),
])
def test_shape_query_matches_oracle(pkg, cfg):
    from oracle import sg3_oracle as orc
    fuw, fuh = cfg['fu']
    fdw, fdh = cfg['fd']
    fu = np.zeros((fuh, fuw) if fuh else (fuw,), np.float32)
    fd = np.zeros((fdh, fdw) if fdh else (fdw,), np.float32)
    oh, ow, sh, swb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = pkg.capi.lib().sg3_filtered_lrelu_shape(cfg['inH'], cfg['inW'], cfg['up'], cfg['down'], fuw, fuh, fdw, fdh,
                                                 *cfg['pad'], ctypes.byref(oh), ctypes.byref(ow), ctypes.byref(sh), ctypes.byref(swb))
    assert rc == 0
    eh, ew = orc.filtered_lrelu_out_shape(cfg['inH'], cfg['inW'], fu, fd, cfg['up'], cfg['down'], cfg['pad'])
    assert (oh.value, ow.value) == (eh, ew)
    assert (sh.value, swb.value) == orc.sign_shape(eh, ew, fd, cfg['down'])
    fu_t = torch.from_numpy(fu)
    fd_t = torch.from_numpy(fd)
    assert pkg.filtered_lrelu.output_shape(cfg['inH'], cfg['inW'], fu_t, fd_t, cfg['up'], cfg['down'], cfg['pad']) == (eh, ew)


def test_shape_query_rejects_tiny_buffer(pkg):
    rc = pkg.capi.lib().sg3_filtered_lrelu_shape(4, 4, 2, 2, 12, 0, 12, 0, 0, 0, 0, 0, None, None, None, None)
    assert rc == pkg.capi.SG3_E_INVALID


def test_setup_filter_contract(pkg):
    up = pkg.upfirdn2d
    f = up.setup_filter([1, 3, 3, 1])
    assert f.shape == (4, 4) and abs(float(f.sum()) - 1) < 1e-6          # short filters become dense outer products
    f8 = up.setup_filter(np.arange(1, 9))
    assert f8.shape == (8,) and abs(float(f8.sum()) - 1) < 1e-6          # >= 8 taps stay separable
    assert up.setup_filter(None).shape == (1, 1)
    g = up.setup_filter([1, 2, 1], gain=4, flip_filter=True)
    assert abs(float(g.sum()) - 4) < 1e-5
    assert up._padding(3) == (3, 3, 3, 3) and up._padding([1, 2]) == (1, 1, 2, 2) and up._padding([1, 2, 3, 4]) == (1, 2, 3, 4)


def test_host_taps_cache_tracks_versions(pkg):
    up = pkg.upfirdn2d
    f = torch.arange(4, dtype=torch.float32)
    a = up.host_taps(f)
    assert up.host_taps(f) is a
    f.mul_(2)
    b = up.host_taps(f)
    assert b is not a and np.allclose(b, [0, 2, 4, 6])


def test_activation_table(pkg):
    t = pkg.bias_act.activation_funcs
    assert list(t) == ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']
    assert [t[k].cuda_idx for k in t] == list(range(1, 10))
    assert t['lrelu'].def_alpha == 0.2 and abs(t['swish'].def_gain - np.sqrt(2)) < 1e-12
    x = torch.randn(5)
    assert torch.allclose(t['lrelu'].func(x, alpha=0.2), torch.nn.functional.leaky_relu(x, 0.2))


def test_cpu_tensors_fail_loudly(pkg):
    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pkg.filtered_lrelu.filtered_lrelu(x)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pkg.upfirdn2d.upfirdn2d(x, None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pkg.bias_act.bias_act(x, torch.zeros(1), act='lrelu')


def test_dropin_seam(pkg):
    saved = {k: v for k, v in sys.modules.items() if k == 'torch_utils' or k.startswith('torch_utils.')}
    try:
        for k in saved:
            del sys.modules[k]
        names = pkg.install()
        assert 'torch_utils.ops.filtered_lrelu' in names
        from torch_utils.ops import filtered_lrelu, bias_act, upfirdn2d, conv2d_gradfix   # the reference's import line
        assert filtered_lrelu is pkg.filtered_lrelu and bias_act is pkg.bias_act and upfirdn2d is pkg.upfirdn2d
        assert conv2d_gradfix.enabled is False
        import inspect
        sig = inspect.signature(filtered_lrelu.filtered_lrelu)
        assert list(sig.parameters) == ['x', 'fu', 'fd', 'b', 'up', 'down', 'padding', 'gain', 'slope', 'clamp', 'flip_filter', 'impl']
        assert list(inspect.signature(bias_act.bias_act).parameters) == ['x', 'b', 'dim', 'act', 'alpha', 'gain', 'clamp', 'impl']
        assert list(inspect.signature(upfirdn2d.upfirdn2d).parameters) == ['x', 'f', 'up', 'down', 'padding', 'flip_filter', 'gain', 'impl']
    finally:
        for k in [k for k in sys.modules if k == 'torch_utils' or k.startswith('torch_utils.')]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_dropin_seam_with_a_real_torch_utils_package(pkg, tmp_path):
    """With a reference-like tree on sys.path, install() must keep the real `torch_utils` package importable
    (networks_stylegan3.py:17 imports torch_utils.misc / persistence next to the aliased ops) and must win over the
    tree's own op modules; patch_modulated_conv() re-points the module global that SynthesisLayer.forward looks up, also
    through a persistence-style wrapper subclass defined in another module."""
    tree = tmp_path / 'fake_ref'
    (tree / 'torch_utils' / 'ops').mkdir(parents=True)
    (tree / 'fakemodels').mkdir()
    (tree / 'torch_utils' / '__init__.py').write_text('')
    (tree / 'torch_utils' / 'misc.py').write_text('MARK = "real misc"\n')
    (tree / 'torch_utils' / 'wrap.py').write_text(
        'def persistent_class(c):\n'
        '    class Decorator(c):\n        pass\n'
        '    Decorator.__name__ = c.__name__\n    return Decorator\n')
    (tree / 'torch_utils' / 'ops' / '__init__.py').write_text('')
    (tree / 'torch_utils' / 'ops' / 'filtered_lrelu.py').write_text('raise ImportError("the tree\'s own plugin wrapper must never be imported")\n')
    (tree / 'fakemodels' / '__init__.py').write_text('')
    (tree / 'fakemodels' / 'net.py').write_text(
        'import torch\nfrom torch_utils import misc, wrap\nfrom torch_utils.ops import filtered_lrelu, bias_act\n'
        'def modulated_conv2d(x, w, s, demodulate=True, padding=0, input_gain=None):\n    return "reference conv"\n'
        '@wrap.persistent_class\nclass Layer(torch.nn.Module):\n'
        '    def forward(self):\n        return modulated_conv2d\n')
    saved = {k: v for k, v in sys.modules.items() if k == 'torch_utils' or k.startswith('torch_utils.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(tree))
    try:
        pkg.install()
        import importlib
        net = importlib.import_module('fakemodels.net')
        assert net.misc.MARK == 'real misc'
        assert net.filtered_lrelu is pkg.filtered_lrelu and net.bias_act is pkg.bias_act
        layer = net.Layer()
        assert type(layer).__module__ == 'torch_utils.wrap'          # like torch_utils.persistence's Decorator subclass
        assert layer.forward()(None, None, None) == 'reference conv'
        flag0 = pkg.filtered_lrelu.round_for_tf32_convs
        assert pkg.patch_modulated_conv(layer) == ['fakemodels.net']
        assert pkg.filtered_lrelu.round_for_tf32_convs is True       # the patched convs read activations with TF32 tensor cores
        pkg.filtered_lrelu.round_for_tf32_convs = flag0              # process-wide switch: leave it as it was for the other tests
        from sg3_b200.modulated_conv import modulated_conv2d
        assert layer.forward() is modulated_conv2d
        assert pkg.patch_modulated_conv(layer) == []                 # idempotent
        assert net._sg3_b200_original_modulated_conv2d(None, None, None) == 'reference conv'
        assert pkg.patch_modulated_conv() == []                      # the reference module itself is not imported here
        with pytest.raises(TypeError):
            pkg.patch_modulated_conv(42)
    finally:
        sys.path.remove(str(tree))
        for k in [k for k in sys.modules if k == 'torch_utils' or k.startswith('torch_utils.') or k.startswith('fakemodels')]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_fov_view_transforms(pkg):
    """sg3_b200.fov.view_transforms vs the construction of utils/fov_expansion.py:34-84 restated here:
    inverse of make_transform(translate, 0) (utils/common.py:9-19) per edge / corner, None for unused views."""
    from sg3_b200 import fov

    def make_transform(translate, angle=0.0):
        m = np.eye(3)
        s, c = np.sin(angle / 360.0 * np.pi * 2), np.cos(angle / 360.0 * np.pi * 2)
        m[0][0], m[0][1], m[0][2], m[1][0], m[1][1], m[1][2] = c, s, translate[0], -s, c, translate[1]
        return m

    res = 1024
    for r, l, t, b in ((100, 50, 30, 70), (0, 64, 0, 0), (10, 0, 0, 20), (0, 0, 0, 0)):
        want = [make_transform((0, 0)),
                make_transform((l / res, 0)) if l else None, make_transform((0, t / res)) if t else None,
                make_transform((-r / res, 0)) if r else None, make_transform((0, -b / res)) if b else None,
                make_transform((l / res, t / res)) if l and t else None, make_transform((-r / res, t / res)) if r and t else None,
                make_transform((-r / res, -b / res)) if r and b else None, make_transform((l / res, -b / res)) if l and b else None]
        got = fov.view_transforms(res, pixels_right=r, pixels_left=l, pixels_top=t, pixels_bottom=b)
        assert len(got) == 9
        for w, g in zip(want, got):
            assert (w is None) == (g is None)
            if w is not None:
                assert np.allclose(np.linalg.inv(w), g, atol=1e-12)
    import inspect
    assert list(inspect.signature(fov.Expander.generate_expanded_image).parameters)[:8] == [
        'self', 'ws', 'all_s', 'landmark_t', 'pixels_right', 'pixels_left', 'pixels_top', 'pixels_bottom']


def test_fov_expander_paste_logic_with_a_stub_generator(pkg):
    """Expander's batching and paste table on a stub generator (CPU tensors; no kernel involved): the stub renders, for every
    sample, an image that encodes which view transform it was given and the pixel coordinates, so the expected canvas follows
    from the paste rules of utils/fov_expansion.py:88-110 alone."""
    from sg3_b200 import fov
    res, n = 16, 2

    class StubInput(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.register_buffer('transform', torch.eye(3))

    class StubSynthesis(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.input = StubInput()
            self.calls = []

        def forward(self, ws, all_s=None, **kw):
            t = self.input.transform
            self.calls.append((ws.shape[0], tuple(t.shape)))
            assert t.shape == (ws.shape[0], 3, 3)
            yy, xx = torch.meshgrid(torch.arange(res), torch.arange(res), indexing='ij')
            base = (yy * res + xx).float()
            # translation in pixels identifies the view; the latent's first entry identifies the sample
            tag = (t[:, 0, 2] * res).round() * 1000 + (t[:, 1, 2] * res).round() * 100000 + ws[:, 0, 0] * 1e7
            return (base[None, None] + tag[:, None, None, None]).expand(-1, 3, -1, -1).contiguous()

    class StubG:
        img_resolution = res
        synthesis = StubSynthesis()

    G = StubG()
    ws = torch.arange(n).float().reshape(n, 1, 1).expand(n, 4, 8).contiguous()
    R, L, T, B = 3, 2, 4, 1
    out = fov.Expander(G).generate_expanded_image(ws=ws, landmark_t=np.eye(3), pixels_right=R, pixels_left=L, pixels_top=T, pixels_bottom=B)
    assert G.synthesis.calls == [(9 * n, (9 * n, 3, 3))]                      # ONE batched call for the nine views
    assert torch.equal(G.synthesis.input.transform, torch.eye(3))             # user transform restored
    assert out.shape == (n, 3, T + res + B, L + res + R)
    yy, xx = np.meshgrid(np.arange(res), np.arange(res), indexing='ij')
    base = (yy * res + xx).astype(np.float64)

    def view(sample, dx, dy):          # image of the view translated by (dx, dy) pixels (inverse of make_transform((dx/res, dy/res)))
        return base + (-dx) * 1000 + (-dy) * 100000 + sample * 1e7

    for smp in range(n):
        want = np.zeros((T + res + B, L + res + R))
        want[T:T + res, L:L + res] = view(smp, 0, 0)
        want[T:T + res, :L] = view(smp, L, 0)[:, :L]
        want[:T, L:L + res] = view(smp, 0, T)[:T, :]
        want[T:T + res, L + res:] = view(smp, -R, 0)[:, res - R:]
        want[T + res:, L:L + res] = view(smp, 0, -B)[res - B:, :]
        want[:T, :L] = view(smp, L, T)[:T, :L]
        want[:T, res + L:] = view(smp, -R, T)[:T, res - R:]
        want[res + T:, res + L:] = view(smp, -R, -B)[res - B:, res - R:]
        want[res + T:, :L] = view(smp, L, -B)[res - B:, :L]
        for ch in range(3):
            assert np.array_equal(out[smp, ch].numpy().astype(np.float64), want), (smp, ch)
    # only the views that are needed are rendered
    G.synthesis.calls.clear()
    out = fov.Expander(G).generate_expanded_image(ws=ws, landmark_t=np.eye(3), pixels_left=5)
    assert G.synthesis.calls == [(2 * n, (2 * n, 3, 3))] and out.shape == (n, 3, res, res + 5)


def test_bench_work_table_matches_oracle_geometry():
    """bench.py derives its algorithmic bytes / FLOPs from the generator it times; the numbers equal those of the oracle's
    independent layer-geometry restatement (R-1024 4.170 GB and 247.6 GFLOP, T-1024 2.109 GB and 570.4 GFLOP)."""
    import sys
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import torch
    import bench
    import sg3_b200  # noqa: F401
    from sg3_b200 import networks
    from oracle import sg3_oracle as orc
    for cfg, gb, gflop in ((bench.R1024, 4.170, 247.6), (bench.T1024, 2.109, 570.4)):
        torch.manual_seed(0)
        G = networks.Generator(**{**cfg, 'img_resolution': 1024})
        specs = []
        for lname in G.synthesis.layer_names:
            L = getattr(G.synthesis, lname)
            specs.append(dict(name=lname, conv_kernel=int(L.conv_kernel), in_size=int(L.in_size[0]), out_size=int(L.out_size[0]),
                              in_channels=int(L.in_channels), out_channels=int(L.out_channels), up=int(L.up_factor),
                              up_taps=int(L.up_taps), padding=list(L.padding),
                              down_filter=None if L.down_filter is None else L.down_filter.numpy()))
        mine = bench.layer_work(specs)
        _, osp = orc.layer_specs(1024, **{k: v for k, v in cfg.items() if k in ('channel_base', 'channel_max', 'conv_kernel', 'use_radial_filters')})
        ref = bench.layer_work(osp)
        assert [(r['flrelu_bytes'], r['conv_flops'], r['flrelu_fma']) for r in mine] == [(r['flrelu_bytes'], r['conv_flops'], r['flrelu_fma']) for r in ref]
        assert abs(sum(r['flrelu_bytes'] for r in mine) / 1e9 - gb) < 2e-3
        assert abs(sum(r['conv_flops'] for r in mine) / 1e9 - gflop) < 0.1


def test_dropin_conv2d_gradfix_stub_has_no_weight_gradients():
    """reference code (setgan/loss.py:152) enters conv2d_gradfix.no_weight_gradients() after install()."""
    from sg3_b200 import dropin
    m = dropin._conv2d_gradfix_module()
    assert m.weight_gradients_disabled is False
    with m.no_weight_gradients():
        assert m.weight_gradients_disabled is True
        with m.no_weight_gradients(False):
            assert m.weight_gradients_disabled is True
    assert m.weight_gradients_disabled is False
    import torch
    y = m.conv2d(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 3, 3), padding=1)
    assert float(y[0, 0, 1, 1]) == 9.0


def test_host_taps_cache_inference_tensors_and_versions():
    """filter buffers created under torch.inference_mode() have no version counter."""
    import torch
    from sg3_b200 import upfirdn2d
    with torch.inference_mode():
        f = torch.arange(4, dtype=torch.float32)
    a = upfirdn2d.host_taps(f)
    assert a.tolist() == [0, 1, 2, 3] and upfirdn2d.host_taps(f) is a
    g = torch.ones(3)
    b = upfirdn2d.host_taps(g)
    assert upfirdn2d.host_taps(g) is b
    g.mul_(2)                                   # in-place update bumps the version: the cache must refresh
    assert upfirdn2d.host_taps(g).tolist() == [2, 2, 2]


def test_build_info_carries_the_source_hash():
    """whether build() recompiled anything must be visible -- the library names the sources it was built from."""
    import importlib.util
    import os
    from sg3_b200 import capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('sg3_b200_build', os.path.join(root, 'stylegan3-editing_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    info = capi.lib().sg3_build_info().decode()
    assert info.endswith('src ' + mod.source_hash()), (info, mod.source_hash())


def test_bench_reference_arm_config_and_workload_harness_import():
    """bench.py's reference arm must carry the GPU arm's config keys; the configs[3] / configs[4] harness models build on CPU."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module('bench')
    cfg = bench.workload_config(1, 32, 'tf32')
    assert set(cfg) >= {'workload', 'per_gpu_batch', 'global_batch', 'parallelism', 'conv_math', 'l2'}
    import torch
    from examples import workloads
    enc = workloads.RestyleEncoder(n_styles=16, input_nc=6).eval()
    with torch.no_grad():
        codes = enc(torch.zeros(1, 6, 256, 256))
    assert tuple(codes.shape) == (1, 16, 512)
    lp = workloads.LPIPSAlex().eval()
    a = torch.rand(1, 3, 64, 64, requires_grad=True)
    d = lp(a, torch.rand(1, 3, 64, 64))
    d.backward()
    assert d.ndim == 0 and a.grad is not None and float(lp(a.detach(), a.detach())) == 0.0
    m = workloads.random_landmarks_transforms(3, torch.Generator().manual_seed(0), 'cpu')
    assert tuple(m.shape) == (3, 3, 3)


def test_rank1_factoring_of_dense_filters():
    """upfirdn2d._rank1_factors: outer-product filters (what setup_filter makes of short 1-D filters) are factored back into their
    1-D taps so that the separable kernel runs them; anything else stays dense."""
    import sg3_b200
    ufd = sg3_b200.upfirdn2d
    f = ufd.setup_filter([1, 3, 3, 1])                       # dense 4 x 4, normalised
    assert f.ndim == 2
    tx, ty = ufd._rank1_factors(ufd.host_taps(f))
    assert np.allclose(np.outer(ty, tx), f.numpy(), rtol=0, atol=1e-7)
    g = np.outer([1.0, -2.0, 0.5], [0.25, 4.0, 1.0, -3.0, 2.0]).astype(np.float32)       # asymmetric shapes and signs
    tx, ty = ufd._rank1_factors(g)
    assert tx.shape == (5,) and ty.shape == (3,) and np.allclose(np.outer(ty, tx), g, atol=1e-6)
    rng = np.random.RandomState(0)
    assert ufd._rank1_factors(rng.randn(4, 4).astype(np.float32)) is None             # full rank
    assert ufd._rank1_factors(np.zeros((3, 3), np.float32)) is None                   # no pivot
    assert ufd._rank1_factors(np.ones((1, 5), np.float32)) is None                    # 1-D shaped filters take their own path
    near = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32)
    near[0, 0] += 1e-3                                                                  # almost rank 1 is not rank 1
    assert ufd._rank1_factors(near) is None


def test_tf32_rounding_switch_logic():
    """filtered_lrelu.tf32_rounded_outputs / round_for_tf32_convs: pure host state (the kernel side is covered on the GPU)."""
    import sg3_b200
    from sg3_b200 import modulated_conv
    fl = sg3_b200.filtered_lrelu
    old = fl.round_for_tf32_convs
    fl.round_for_tf32_convs = False                           # an earlier drop-in test may have patched modulated_conv2d
    assert fl._rounding_wanted() is False
    with fl.tf32_rounded_outputs(True):
        assert fl._rounding_wanted() is True
        with fl.tf32_rounded_outputs(False):
            assert fl._rounding_wanted() is False
        assert fl._rounding_wanted() is True
    assert fl._rounding_wanted() is False
    try:
        fl.round_for_tf32_convs = True                        # what patch_modulated_conv() sets
        modulated_conv.set_math('tf32')
        assert modulated_conv.tf32_activation_policy() == 'compensate'
        assert fl._rounding_wanted() is False                 # default policy: the conv weights carry the correction instead
        modulated_conv.set_tf32_activation_policy('round')
        assert fl._rounding_wanted() is True
        modulated_conv.set_math('fp32')
        assert fl._rounding_wanted() is False
        modulated_conv.set_math('fp32x3')
        assert fl._rounding_wanted() is False
    finally:
        fl.round_for_tf32_convs = old
        modulated_conv.set_tf32_activation_policy('compensate')
        modulated_conv.set_math(None)
    assert sg3_b200.capi.FLRELU_ROUND_TF32 == 1


def test_row_pitched_helpers():
    """Host logic of the TMA hand-over buffers: `empty_row_pitched` pads rows to 16 bytes (4 floats / 8 halves) and zeroes the padding,
    `_row_pitched` recognises exactly those views, `_tma_rows` copies only what TMA cannot address."""
    import torch
    from sg3_b200 import modulated_conv as mc
    t = mc.empty_row_pitched([2, 3, 5, 38], torch.float32, 'cpu')
    assert tuple(t.shape) == (2, 3, 5, 38) and t.stride() == (3 * 5 * 40, 5 * 40, 40, 1) and mc._row_pitched(t)
    assert float(t.as_strided((2, 3, 5, 2), t.stride(), 38).abs().max()) == 0.0          # the padding columns
    h = mc.empty_row_pitched([1, 2, 4, 1044], torch.float16, 'cpu', align=8)
    assert h.stride(2) == 1048 and mc._row_pitched(h)
    dense = torch.zeros(1, 2, 4, 40)
    assert mc._row_pitched(dense) and mc._tma_rows(dense) is dense
    assert not mc._row_pitched(dense.transpose(2, 3))
    odd = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 3, 4, 6)            # width 6: rows are 24 bytes apart
    p = mc._tma_rows(odd)
    assert p is not odd and p.stride(2) == 8 and torch.equal(p, odd)


def test_conv_backward_entry_points_validate_before_launching(pkg):
    """Argument checks of the new C entry points run on the host before any CUDA call: invalid -> SG3_E_INVALID, shapes / pitches
    without a kernel -> SG3_E_NOKERNEL (the Python op then falls back), never a launch with bad geometry."""
    L = pkg.capi.lib()
    E_INVALID, E_NOKERNEL = pkg.capi.SG3_E_INVALID, pkg.capi.SG3_E_NOKERNEL
    a = 4096                                                    # any non-null, 16-byte aligned address: nothing is dereferenced
    assert L.sg3_modconv_wgrad3(None, a, a, 1, 8, 8, 10, 12, 2, 8, 16, 0, None) == E_INVALID
    assert L.sg3_modconv_wgrad3(a, a, a, 1, 8, 8, 10, 12, 2, 4, 16, 0, None) == E_INVALID        # ldw < I
    assert L.sg3_modconv_wgrad3(a, a, a, 1, 8, 8, 10, 12, 1, 8, 16, 0, None) == E_NOKERNEL       # padding 1
    assert L.sg3_modconv_wgrad3(a, a, a, 1, 8, 8, 10, 12, 2, 8, 0, 0, None) == E_NOKERNEL        # OW = 14: not a 16-byte pitch
    assert L.sg3_modconv_wgrad3(a, a + 4, a, 1, 8, 8, 10, 12, 2, 8, 16, 0, None) == E_NOKERNEL   # x off the TMA alignment
    assert L.sg3_modconv_wgrad3(a, a, a + 8, 1, 8, 8, 10, 12, 2, 8, 16, 0, None) == E_NOKERNEL   # dw off the 128-bit reduction alignment
    assert L.sg3_modconv_weights_bwd_taps(a, a, a, None, 0, a, a, a, 1, 513, 8, 3, 516, 1, None) == E_NOKERNEL   # > 512 channels x 9 taps
    assert L.sg3_modconv_weights_bwd_taps(a, a, a, None, 1, a, a, a, 1, 8, 8, 3, 8, 1, None) == E_INVALID        # gain mode without a gain
    assert L.sg3_modconv_weights(a, a, None, 0, a, a, 1, 8, 8, 3, 72, 1, 3, 2, None) == E_INVALID                 # 3xTF32 planes are layout 0 only
    assert L.sg3_modconv_weights(a, a, None, 0, a, a, 1, 8, 8, 3, 72, 1, 2, 3, None) == E_INVALID                 # fp16 has no dgrad-tap layout
