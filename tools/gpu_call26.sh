#!/bin/bash
# GPU call 26 of round 2: ncu of the two depthwise 7x7 kernels at the P3 shape.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dwconv -o gpurun_out/dwconv_r02y python tools/profile_dwconv.py > gpurun_out/dwconv_ncu_y.log 2>&1
tail -3 gpurun_out/dwconv_ncu_y.log
