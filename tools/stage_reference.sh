#!/bin/sh
# Stage the UNMODIFIED reference files the GPU-side reference checks need into baseline/_ref/ (git-ignored, travels to
# the GPU box with the gpurun snapshot).  Run in the authoring container, where /root/reference exists.  Nothing is edited.
#   * models/stylegan3 + torch_utils/{misc,persistence} + dnnlib   -> tools/check_dropin_reference.py
#   * torch_utils/custom_ops.py + torch_utils/ops/*                 -> tools/check_vs_reference_cuda.py, and
#     the three CUDA plugins (bias_act_plugin, upfirdn2d_plugin, filtered_lrelu_plugin) PREBUILT here from those
#     sources with the reference's own flags (custom_ops.get_plugin: --use_fast_math) for sm_100 into
#     baseline/_ref/_plugins/, so the GPU box does not spend ~3 min of every call JIT-compiling them.
set -e
SRC=${SG3_REF_ROOT:-/root/reference}
DST=$(dirname "$0")/../baseline/_ref
mkdir -p "$DST/models/stylegan3" "$DST/torch_utils/ops"
cp -r "$SRC/dnnlib" "$DST/"
cp "$SRC/torch_utils/__init__.py" "$SRC/torch_utils/misc.py" "$SRC/torch_utils/persistence.py" "$SRC/torch_utils/custom_ops.py" "$DST/torch_utils/"
cp "$SRC"/torch_utils/ops/*.py "$SRC"/torch_utils/ops/*.cpp "$SRC"/torch_utils/ops/*.cu "$SRC"/torch_utils/ops/*.h "$DST/torch_utils/ops/"
cp "$SRC/models/__init__.py" "$DST/models/" 2>/dev/null || true
cp "$SRC/models/stylegan3/__init__.py" "$SRC/models/stylegan3/networks_stylegan3.py" "$DST/models/stylegan3/"
echo "staged into $DST"
# A generator pickled by the reference's own persistence machinery (the pickle embeds the reference's module source, so it
# lives in the git-ignored staging area too): tiny R, seed 0 -- the weights of tests/golden/tiny.npz.
python - "$SRC" "$DST" <<'PY'
import pickle, sys
sys.path.insert(0, sys.argv[1])
import torch
from models.stylegan3 import networks_stylegan3 as ref
torch.manual_seed(0)
G = ref.Generator(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
                  conv_kernel=1, use_radial_filters=True).eval().requires_grad_(False)
pickle.dump(dict(G_ema=G), open(sys.argv[2] + '/tinyR_seed0.pkl', 'wb'))
print('pickled', sys.argv[2] + '/tinyR_seed0.pkl')
PY
# The reference's CUDA plugins, compiled from the staged (unmodified) sources exactly as custom_ops.get_plugin would on a
# B200 (torch.utils.cpp_extension.load, --use_fast_math; arch 10.0+PTX is what torch derives from the device there).
if [ "${SG3_SKIP_PLUGINS:-0}" != "1" ]; then
python - "$DST" <<'PY'
import os, sys, time
os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0+PTX'
import torch.utils.cpp_extension as ce
dst = os.path.abspath(sys.argv[1])
ops = os.path.join(dst, 'torch_utils', 'ops')
plugins = {
    'bias_act_plugin': ['bias_act.cpp', 'bias_act.cu'],
    'upfirdn2d_plugin': ['upfirdn2d.cpp', 'upfirdn2d.cu'],
    'filtered_lrelu_plugin': ['filtered_lrelu.cpp', 'filtered_lrelu_wr.cu', 'filtered_lrelu_rd.cu', 'filtered_lrelu_ns.cu'],
}
for name, srcs in plugins.items():
    bd = os.path.join(dst, '_plugins', name)
    os.makedirs(bd, exist_ok=True)
    t0 = time.time()
    ce.load(name=name, sources=[os.path.join(ops, s) for s in srcs], build_directory=bd, verbose=False,
            extra_cuda_cflags=['--use_fast_math', '--allow-unsupported-compiler'], is_python_module=False)
    for f in os.listdir(bd):                      # keep the .so only (the objects are 30 MB of dead weight in the snapshot)
        if f.endswith('.o'):
            os.remove(os.path.join(bd, f))
    print(f'built {name} in {time.time() - t0:.0f} s ->', [f for f in os.listdir(bd) if f.endswith('.so')])
PY
fi
