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
