#!/bin/bash
# GPU call 22 of round 2 (2 GPUs): the N=2 bench line exactly as the driver launches it, then the N=1 line (predictor leg with cached staging).
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02u_n2.json 2> gpurun_out/bench_r02u_n2.err
tail -c 600 gpurun_out/bench_r02u_n2.json | head -c 300; echo
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02u_n2.json"))
    print("N=2", d["value"], d["n_gpus"], d["e2e"]["value"], d["ms_per_step"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02u_n2.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02u_n1.json 2> gpurun_out/bench_r02u_n1.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02u_n1.json"))
    print("N=1", d["value"], d["e2e"]["value"], d["single_stream"]["value"], d["e2e_predictor"])
except Exception as e: print("bench ERR", e)
PY
