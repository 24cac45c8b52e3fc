#!/bin/bash
# GPU call 18 of round 2: launch list of the current forward, throughput of the other model families, ncu of the MoE kernels.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02r.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02r.log 2>&1
timeout 900 python tools/bench_models.py 32 gpurun_out/models_r02r.json > gpurun_out/models_r02r.log 2>&1; tail -15 gpurun_out/models_r02r.log | cut -c1-200
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"moe_ffn|moe_combine|router_fused" -c 5 -o gpurun_out/moe_r02r python tools/profile_forward.py > gpurun_out/moe_ncu_r.log 2>&1
