"""Drop-in seam: expose this package's op modules as `torch_utils.ops.{filtered_lrelu,bias_act,
upfirdn2d,conv2d_gradfix}` so the reference's model code (`from torch_utils.ops import ...`,
networks_stylegan3.py:18) and pickled generators pick them up unchanged."""
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

    m.conv2d = conv2d
    m.conv_transpose2d = conv_transpose2d
    return m


def install(override_existing=True):
    from . import bias_act, filtered_lrelu, upfirdn2d
    mods = {
        'torch_utils.ops.filtered_lrelu': filtered_lrelu,
        'torch_utils.ops.bias_act': bias_act,
        'torch_utils.ops.upfirdn2d': upfirdn2d,
    }
    if 'torch_utils' not in sys.modules:
        pkg = types.ModuleType('torch_utils')
        pkg.__path__ = []
        sys.modules['torch_utils'] = pkg
    if 'torch_utils.ops' not in sys.modules:
        ops = types.ModuleType('torch_utils.ops')
        ops.__path__ = []
        sys.modules['torch_utils.ops'] = ops
        setattr(sys.modules['torch_utils'], 'ops', ops)
    ops = sys.modules['torch_utils.ops']
    if 'torch_utils.ops.conv2d_gradfix' not in sys.modules:
        mods['torch_utils.ops.conv2d_gradfix'] = _conv2d_gradfix_module()
    for name, mod in mods.items():
        if override_existing or name not in sys.modules:
            sys.modules[name] = mod
            setattr(ops, name.rsplit('.', 1)[1], mod)
    return sorted(mods)
