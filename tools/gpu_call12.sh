#!/bin/bash
# GPU call 12 of round 2: ncu --set full of the tc_conv2 family (first 12 launches of the forward), the 7x7 depthwise and the fused router.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_conv2" -c 12 -o gpurun_out/tcconv2_r02l python tools/profile_forward.py > gpurun_out/tcconv2_ncu_l.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dwconv_tiled|router_fused" -c 3 -o gpurun_out/dw_router_r02l python tools/profile_forward.py > gpurun_out/dw_router_ncu_l.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
