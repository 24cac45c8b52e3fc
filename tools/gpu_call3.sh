#!/bin/bash
# GPU call 3 of round 2: microbenchmarks, suite, PDL A/B, attention2 ncu, reference arm at bs32, racecheck.
mkdir -p gpurun_out
timeout 120 ./tools/ubench_tmem.bin > gpurun_out/ubench_tmem.txt 2>&1
bash tools/gpu_suite.sh r02c
YM_PDL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02c_pdl1.json 2> gpurun_out/bench_r02c_pdl1.err
YM_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02c_pdl0.json 2> gpurun_out/bench_r02c_pdl0.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_attention2 -c 1 -o gpurun_out/attn2_r02c python tools/profile_attention.py > gpurun_out/attn2_ncu.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r02c.json 2> gpurun_out/bench_ref_r02c.err
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_tc.py -q -x -p no:cacheprovider -k "attention2_strict and (400-2-32-3 or 100-2-64-1)" > gpurun_out/racecheck_attention2.log 2>&1
python - <<'PY'
import json
for t in ("pdl1", "pdl0"):
    try:
        d = json.load(open(f"gpurun_out/bench_r02c_{t}.json")); print(t, d["value"], d["e2e"]["value"], d["kernels_per_step"], d["roofline"]["ms_per_launch"], d["dispatch"]["hbm_frac"])
    except Exception as e: print(t, "ERR", e)
PY
cat gpurun_out/ubench_tmem.txt; head -c 300 gpurun_out/bench_ref_r02c.json; tail -5 gpurun_out/racecheck_attention2.log
