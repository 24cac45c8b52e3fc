#!/bin/bash
# GPU call 2 of round 2: whole suite, attention A/B timing, cls diagnostic, bench (both arms).
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
bash tools/gpu_suite.sh r02b
timeout 300 python tools/attn_bench.py gpurun_out/attn_bench_r02b.json > gpurun_out/attn_bench_r02b.log 2>&1
timeout 300 python tools/diag_cls.py > gpurun_out/diag_cls.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
avail=$(free -g | awk '/Mem:/{print $7}')
if [ "$avail" -gt 128 ]; then
  /usr/bin/time -v timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r02b.json 2> gpurun_out/bench_ref_r02b.err
fi
tail -3 gpurun_out/attn_bench_r02b.log; tail -2 gpurun_out/diag_cls.log; head -c 600 gpurun_out/bench_r02b.json; echo; head -c 400 gpurun_out/bench_ref_r02b.json
