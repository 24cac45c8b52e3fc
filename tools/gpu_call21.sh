#!/bin/bash
# GPU call 21 of round 2: bench at the new default depth (4 graph instances), twice (run-to-run spread).
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02t$i.json 2> gpurun_out/bench_r02t$i.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r02t$i.json"))
    print("bench", d["value"], d["e2e"]["value"], d["e2e"]["unpipelined_value"], d["single_stream"]["value"], d["e2e_predictor"]["value"])
except Exception as e: print("bench ERR", e)
PY
done
