#!/bin/bash
# GPU call 13 of round 2: tc_conv2 epilogue-store variant, coalesced fp32 logits, rowmax8; ops tests + op bench + bench + launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_model_v0.py tests/test_gpu_pins.py -q -m gpu > gpurun_out/tests_r02m.log 2>&1; tail -6 gpurun_out/tests_r02m.log
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02m.json > gpurun_out/op_bench_r02m.log 2>&1; grep -E "^conv |Error|error" gpurun_out/op_bench_r02m.log | cut -c1-150
timeout 900 python bench.py > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02m.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02m.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02m.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02m.err
