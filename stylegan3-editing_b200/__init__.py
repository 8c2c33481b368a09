"""sg3_b200 -- B200-native (sm_100a) kernels for the StyleGAN3 synthesis hot path of
krylea/stylegan3-editing, behind the reference's own op API.

    import sg3_b200                         # repo-root shim; this directory's name has a hyphen
    from sg3_b200 import filtered_lrelu, bias_act, upfirdn2d
    y = filtered_lrelu.filtered_lrelu(x, fu, fd, b, up=2, down=2, padding=[11, 10, 11, 10], clamp=256)

    sg3_b200.install()                      # alias the ops as torch_utils.ops.* for the reference's models

    from sg3_b200.fov import Expander       # utils/fov_expansion.Expander in one batched synthesis call

Everything is backed by libsg3_b200.so (C ABI in include/sg3_b200.h); there is no CPU fallback.
"""
from . import capi
from . import bias_act, filtered_lrelu, upfirdn2d   # noqa: F401  (reference-compatible op modules)

__all__ = ['capi', 'bias_act', 'filtered_lrelu', 'upfirdn2d', 'install', 'patch_modulated_conv']


def install():
    """Register the op modules under the names the reference imports (`torch_utils.ops.*`).

    Call before importing the reference's `models.stylegan3.networks_stylegan3` (or unpickling a
    generator: `torch_utils/persistence.py` re-imports ops by module name), and the unmodified
    reference model code runs on these kernels.
    """
    from . import dropin
    return dropin.install()


def patch_modulated_conv(target=None, round_activations=True):
    """Also route the reference's module-level `modulated_conv2d` (pure PyTorch + cuDNN, networks_stylegan3.py:24-63) to the
    fused prologue + tcgen05 contraction; `target` = None (the imported `models.stylegan3.networks_stylegan3`), a module, or a
    generator object (covers unpickled generators).  See dropin.patch_modulated_conv."""
    from . import dropin
    return dropin.patch_modulated_conv(target, round_activations=round_activations)
