"""Drop-in seam: expose this package's op modules as `torch_utils.ops.{filtered_lrelu,bias_act,
upfirdn2d,conv2d_gradfix}` so the reference's model code (`from torch_utils.ops import ...`,
networks_stylegan3.py:18) and pickled generators pick them up unchanged."""
import contextlib
import importlib
import sys
import types

import torch


def _conv2d_gradfix_module():
    """Minimal stand-in for torch_utils/ops/conv2d_gradfix.py: `enabled` is False by default in the
    reference (:22), in which case conv2d/conv_transpose2d are plain F.conv2d / F.conv_transpose2d."""
    m = types.ModuleType('torch_utils.ops.conv2d_gradfix')
    m.enabled = False
    m.weight_gradients_disabled = False

    def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                          dilation=dilation, groups=groups)

    def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
        return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                    output_padding=output_padding, groups=groups, dilation=dilation)

    @contextlib.contextmanager
    def no_weight_gradients(disable=True):
        # conv2d_gradfix.py:25-33: a flag the reference's custom conv backward consults; with `enabled = False` (plain
        # F.conv2d) it changes nothing, but callers such as setgan/loss.py:152 still enter the context
        old = m.weight_gradients_disabled
        if disable:
            m.weight_gradients_disabled = True
        yield
        m.weight_gradients_disabled = old

    m.conv2d = conv2d
    m.conv_transpose2d = conv_transpose2d
    m.no_weight_gradients = no_weight_gradients
    return m


def install(override_existing=True):
    from . import bias_act, filtered_lrelu, upfirdn2d
    mods = {
        'torch_utils.ops.filtered_lrelu': filtered_lrelu,
        'torch_utils.ops.bias_act': bias_act,
        'torch_utils.ops.upfirdn2d': upfirdn2d,
    }
    # Prefer the real packages when the reference tree is on sys.path (its `torch_utils.misc` / `.persistence` must stay
    # importable: networks_stylegan3.py:17); synthesise empty ones only when there is no `torch_utils` at all.  Importing the
    # reference's `torch_utils.ops` package itself is harmless (empty __init__); its op submodules are never imported because
    # the aliases below are registered first.
    for name in ('torch_utils', 'torch_utils.ops'):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                pkg = types.ModuleType(name)
                pkg.__path__ = []
                sys.modules[name] = pkg
    if getattr(sys.modules['torch_utils'], 'ops', None) is not sys.modules['torch_utils.ops']:
        setattr(sys.modules['torch_utils'], 'ops', sys.modules['torch_utils.ops'])
    ops = sys.modules['torch_utils.ops']
    if 'torch_utils.ops.conv2d_gradfix' not in sys.modules:
        # the reference's own conv2d_gradfix (pure Python, no plugin) when its tree is importable; the stand-in otherwise
        try:
            importlib.import_module('torch_utils.ops.conv2d_gradfix')
        except Exception:
            mods['torch_utils.ops.conv2d_gradfix'] = _conv2d_gradfix_module()
    for name, mod in mods.items():
        if override_existing or name not in sys.modules:
            sys.modules[name] = mod
            setattr(ops, name.rsplit('.', 1)[1], mod)
    return sorted(mods)


def patch_modulated_conv(target=None, round_activations=True):
    """Point the reference's module-level `modulated_conv2d` (networks_stylegan3.py:24, looked up as a module global by
    `SynthesisLayer.forward`, :360) at this package's fused implementation (same signature).

    `target` may be
      * None: `models.stylegan3.networks_stylegan3`, if it has been imported;
      * a module;
      * a generator / any `torch.nn.Module`: every Python module that defines the class of one of its submodules and has a
        `modulated_conv2d` global is patched -- this covers pickled generators, whose source `torch_utils/persistence.py:191-229`
        re-imports under a private module name.
    `round_activations`: once a module is patched, `filtered_lrelu.round_for_tf32_convs` is switched on -- in TF32 math mode every fused
    filtered_lrelu forward then rounds its fp32 outputs to the nearest TF32 value, because the patched convolutions read them with
    tensor cores that would truncate (2-3x lower image error, see filtered_lrelu.tf32_rounded_outputs).  It is a process-wide switch:
    pass False (or reset the attribute) if the same process also calls filtered_lrelu for consumers that want unrounded fp32 values.
    Returns the names of the patched modules (empty list: nothing to patch)."""
    from .modulated_conv import modulated_conv2d
    mods = []
    if target is None:
        m = sys.modules.get('models.stylegan3.networks_stylegan3')
        if m is not None:
            mods.append(m)
    elif isinstance(target, types.ModuleType):
        mods.append(target)
    elif isinstance(target, torch.nn.Module):
        seen = set()
        for sub in target.modules():
            # walk the MRO: `persistence.persistent_class` wraps every reference class in a subclass that lives in
            # torch_utils.persistence; the class whose forward() looks up `modulated_conv2d` is one step up
            for klass in type(sub).__mro__:
                name = klass.__module__
                if name not in seen:
                    seen.add(name)
                    m = sys.modules.get(name)
                    if m is not None:
                        mods.append(m)
    else:
        raise TypeError('patch_modulated_conv: expected None, a module or a torch.nn.Module')
    done = []
    for m in mods:
        if callable(getattr(m, 'modulated_conv2d', None)) and m.modulated_conv2d is not modulated_conv2d:
            m._sg3_b200_original_modulated_conv2d = m.modulated_conv2d
            m.modulated_conv2d = modulated_conv2d
            done.append(m.__name__)
    if done and round_activations:
        # the patched layers' convolutions now read their inputs with TF32 tensor cores (when the math mode is 'tf32'): have the
        # stencils that produce those inputs round them to nearest (filtered_lrelu.tf32_rounded_outputs)
        from . import filtered_lrelu
        filtered_lrelu.round_for_tf32_convs = True
    return done
