#!/bin/bash
# GPU call 9 of round 2: patch-staged small 3x3 kernel, stem mma.sync kernel with fragments from shared memory, concat2; op timings + bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_model_v0.py -q -m gpu > gpurun_out/tests_r02i.log 2>&1; tail -15 gpurun_out/tests_r02i.log
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02i.json > gpurun_out/op_bench_r02i.log 2>&1; cat gpurun_out/op_bench_r02i.log | cut -c1-150
timeout 900 python bench.py > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02i.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02i.err
