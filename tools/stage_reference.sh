#!/bin/sh
# Stage the few UNMODIFIED reference files tools/check_dropin_reference.py needs into baseline/_ref/ (git-ignored, travels to
# the GPU box with the gpurun snapshot).  Run in the authoring container, where /root/reference exists.  Nothing is edited.
set -e
SRC=${SG3_REF_ROOT:-/root/reference}
DST=$(dirname "$0")/../baseline/_ref
mkdir -p "$DST/models/stylegan3" "$DST/torch_utils/ops"
cp -r "$SRC/dnnlib" "$DST/"
cp "$SRC/torch_utils/__init__.py" "$SRC/torch_utils/misc.py" "$SRC/torch_utils/persistence.py" "$DST/torch_utils/"
cp "$SRC/torch_utils/ops/__init__.py" "$DST/torch_utils/ops/"
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
