#!/bin/bash
# GPU call 17 of round 2: persistent conv kernel with shared-space staging stores, division-free tile walk, lighter barrier waits: tests, op bench (groups 2 vs 1), bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_model_v0.py tests/test_gpu_pins.py tests/test_gpu_zz_segment.py tests/test_gpu_zz_pose.py tests/test_gpu_mot.py -q -m gpu -x > gpurun_out/tests_r02q.log 2>&1; tail -6 gpurun_out/tests_r02q.log
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02q.json > gpurun_out/op_bench_r02q.log 2>&1; grep -E "^conv |Error|error|assert" gpurun_out/op_bench_r02q.log | cut -c1-150
timeout 900 python bench.py > gpurun_out/bench_r02q.json 2> gpurun_out/bench_r02q.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02q.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02q.err
