#!/bin/bash
# Scratch GPU call: targeted tests + op bench + bench (edit as needed).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_zz_predictor.py -q -m gpu -x > gpurun_out/tests_probe.log 2>&1; tail -4 gpurun_out/tests_probe.log | cut -c1-200
timeout 600 python tools/op_bench.py gpurun_out/op_bench_probe.json > gpurun_out/op_bench_probe.log 2>&1; grep -E "^dwconv7|Error|error|assert" gpurun_out/op_bench_probe.log | cut -c1-150
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_probe.json 2> gpurun_out/bench_probe.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_probe.json"))
    print("bench", d["value"], d["e2e"]["value"], d["single_stream"]["value"], d["e2e_predictor"])
except Exception as e: print("bench ERR", e)
PY
