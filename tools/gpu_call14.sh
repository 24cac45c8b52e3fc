#!/bin/bash
# GPU call 14 of round 2: tc_conv2 timing probe; bench with the TMA-store default restored and the predictor end-to-end leg.
mkdir -p gpurun_out
timeout 600 python tools/conv2_probe.py gpurun_out/conv2_probe_r02n.json > gpurun_out/conv2_probe_r02n.log 2>&1; cut -c1-150 gpurun_out/conv2_probe_r02n.log
timeout 900 python bench.py > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02n.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"]); print("predictor", d.get("e2e_predictor"))
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02n.err
