#!/bin/bash
# Round-end GPU call: the whole GPU suite (no -x), the driver's two bench arms at N=1, launch list of the final forward.
#   bash tools/gpurun_retry.sh 3000 'bash tools/gpu_final.sh r02z'
tag=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_suite.sh $tag
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"], d["e2e_predictor"].get("value"))
    print("eager", d["torch_eager_gpu"]["value"], "cpu", d["cpu_baseline"]["value"], "disp", d["dispatch"]["hbm_frac"], "attn", d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; head -c 400 gpurun_out/bench_ref_$tag.json; echo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_$tag.csv python tools/profile_forward.py > gpurun_out/profile_forward_$tag.log 2>&1
