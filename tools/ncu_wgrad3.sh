#!/bin/bash
# ncu --set full of one sg3_modconv_wgrad3 launch at a StyleGAN3-T layer shape:  tools/ncu_wgrad3.sh L11 [tag]
L=${1:-L11}; TAG=${2:-r02i}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:modconv_wgrad3 --launch-skip 3 -c 1 \
    -o gpurun_out/${TAG}_wgrad3_${L} -f python tools/prof_wgrad3.py 4 ${L}_ > gpurun_out/${TAG}_ncu_wgrad3_${L}.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_wgrad3_${L}.log
