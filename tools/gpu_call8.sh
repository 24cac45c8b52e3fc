#!/bin/bash
# GPU call 8 of round 2: attention scheduling variants, MoE kernels back at two CTAs per SM, stem staging; targeted tests + bench + launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model_v0.py tests/test_gpu_model.py -q -m gpu -x > gpurun_out/tests_r02h.log 2>&1; tail -4 gpurun_out/tests_r02h.log
timeout 600 python tools/attn_bench.py gpurun_out/attn_bench_r02h.json > gpurun_out/attn_bench_r02h.log 2>&1; grep -E "P3|P4" gpurun_out/attn_bench_r02h.log | grep -E "impl 2|variant" | cut -c1-150
timeout 900 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02h.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02h.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02h.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"])
    print("disp", d["dispatch"]["hbm_frac"], "attn", d["roofline"]["ms_per_launch"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02h.err
grep -E "moe_|stem_conv|gemm_conv_kernel<(32|8|16), 0, 0>" gpurun_out/launch_dram_r02h.csv | awk -F'","' '{print $5, $(NF)}' | head -0
