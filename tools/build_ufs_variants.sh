#!/bin/sh
# Tuning builds of ONE source file: libsg3_b200_<tag>.so = the default objects + $SRC (default upfirdn2d_stream) compiled with
# extra -D flags.   [SRC=modconv_tc] tools/build_ufs_variants.sh tag "-DUFS_SMEM_MIN=46000" [tag2 "..."]
set -e
P=stylegan3-editing_b200
SRC=${SRC:-upfirdn2d_stream}
python $P/build.py > /dev/null
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  mkdir -p $P/csrc/build_$tag
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
    --expt-relaxed-constexpr -I include $flags -c $P/csrc/$SRC.cu -o $P/csrc/build_$tag/$SRC.o
  objs=$(ls $P/csrc/build/*.o | grep -v /$SRC.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/libsg3_b200_$tag.so $objs $P/csrc/build_$tag/$SRC.o -cudart static
  echo built $P/libsg3_b200_$tag.so
done
