#!/bin/bash
# GPU call 11 of round 2: full GPU suite after the stem / small-conv / rowmax / q_tiles changes, op bench, bench.
mkdir -p gpurun_out
bash tools/gpu_suite.sh r02k
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02k.json > gpurun_out/op_bench_r02k.log 2>&1; cut -c1-150 gpurun_out/op_bench_r02k.log
timeout 900 python bench.py > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02k.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02k.err
