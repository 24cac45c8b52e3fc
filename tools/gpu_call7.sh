#!/bin/bash
# GPU call 6 of round 2: suite, bench (PipelinedForward default + depth sweep), launch roofline, ncu of the stem / MoE kernels.
mkdir -p gpurun_out
bash tools/gpu_suite.sh r02g
timeout 900 python bench.py --depth-sweep > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launch_dram_r02g.csv python tools/profile_forward.py > gpurun_out/profile_forward_r02g.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"moe_|stem_conv" -c 5 -o gpurun_out/moe_stem_r02g python tools/profile_forward.py > gpurun_out/moe_stem_ncu_g.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02g.json"))
    print("bench", d["value"], d["e2e"]["value"], d["kernels_per_step"], d["single_stream"], d.get("depth_sweep_images_per_s"))
    print("eager", d["torch_eager_gpu"]); print("cpu", d["cpu_baseline"]); print("disp", d["dispatch"]["hbm_frac"], "attn", d["roofline"]["ms_per_launch"])
except Exception as e: print("bench ERR", e)
PY
tail -3 gpurun_out/bench_r02g.err
