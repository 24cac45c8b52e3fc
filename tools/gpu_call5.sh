#!/bin/bash
# GPU call 4 of round 2: suite (new fused MoE-FFN kernels, two-issuer attention), attention A/B + poly sweep, bench, launch roofline, ncu.
mkdir -p gpurun_out
bash tools/gpu_suite.sh r02e
timeout 300 python tools/attn_bench.py gpurun_out/attn_bench_r02e.json > gpurun_out/attn_bench_r02e.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --two-stream > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02e.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_attention2 -c 1 -o gpurun_out/attn2_r02e python tools/profile_attention.py > gpurun_out/attn2_ncu_e.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:moe_ -c 6 -o gpurun_out/moe_ffn_r02e python tools/profile_forward.py > gpurun_out/moe_ffn_ncu_e.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02e.json")); print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["dispatch"]["hbm_frac"])
except Exception as e: print("bench ERR", e)
PY
grep -E "P3 impl|P3 poly|P4 impl" gpurun_out/attn_bench_r02e.log | cut -c1-200
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02e.json")); print("two_stream", d.get("two_stream"))
PY
