#!/bin/bash
# GPU call 25 of round 2: depthwise 7x7 on mma.sync (Toeplitz GEMM): tests, op bench, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_model_v0.py tests/test_gpu_pins.py -q -m gpu -x > gpurun_out/tests_r02x.log 2>&1; tail -12 gpurun_out/tests_r02x.log | cut -c1-200
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02x.json > gpurun_out/op_bench_r02x.log 2>&1; grep -E "^dwconv7|Error|error|assert" gpurun_out/op_bench_r02x.log | cut -c1-150
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02x.json 2> gpurun_out/bench_r02x.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02x.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02x.err
