#!/bin/sh
# Time the main filtered_lrelu layer shapes with every tuning build present (libsg3_b200_*.so, see build.py).
for lib in stylegan3-editing_b200/libsg3_b200.so stylegan3-editing_b200/libsg3_b200_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  for L in ${LAYERS:-L11 L10 L12 L6 L5}; do
    SG3_B200_LIB=$PWD/$lib python tools/prof_flrelu.py $L 2 2>&1 | grep -v Warn | head -1
  done
done
