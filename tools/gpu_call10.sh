#!/bin/bash
# GPU call 10 of round 2: launch list after the small-conv / stem / concat changes, graph-timed op bench, q_tiles sweep, ncu of the new kernels.
mkdir -p gpurun_out
timeout 600 python tools/op_bench.py gpurun_out/op_bench_r02j.json > gpurun_out/op_bench_r02j.log 2>&1; cut -c1-150 gpurun_out/op_bench_r02j.log
timeout 600 python tools/attn_bench.py gpurun_out/attn_bench_r02j.json > gpurun_out/attn_bench_r02j.log 2>&1; grep -E "q_tiles" gpurun_out/attn_bench_r02j.log | cut -c1-150
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02j.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02j.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"small_conv|stem_conv" -c 4 -o gpurun_out/small_stem_r02j python tools/profile_forward.py > gpurun_out/small_stem_ncu_j.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02j.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"]["value"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02j.err
