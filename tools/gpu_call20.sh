#!/bin/bash
# GPU call 20 of round 2: bench with the depth sweep (router kernel at two CTAs per SM), launch list.
mkdir -p gpurun_out
timeout 900 python bench.py --depth-sweep > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02s.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"], d.get("depth_sweep_images_per_s"))
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02s.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02s.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02s.log 2>&1
grep -c tc_conv2 gpurun_out/launch_dram_r02s.csv
