#!/bin/bash
# GPU call 19 of round 2: compute-sanitizer racecheck on the tcgen05 / TMA kernels (small cases of their strict tests).
mkdir -p gpurun_out
run() {  # $1 tag, $2 pytest file, $3 -k expression
  echo "# compute-sanitizer --tool racecheck python -m pytest $2 -k '$3'" > gpurun_out/racecheck_$1.txt
  timeout 600 compute-sanitizer --tool racecheck python -m pytest $2 -q -m gpu -k "$3" -x -p no:cacheprovider >> gpurun_out/racecheck_$1.txt 2>&1
  tail -3 gpurun_out/racecheck_$1.txt
}
run attention2 tests/test_gpu_tc.py "attention2_strict and (400-2-32-3 or 100-2-64-1 or 257-2-32-2)"
run conv2 tests/test_gpu_tc.py "test_tc_conv_matches_oracle"
run moe tests/test_gpu_tc.py "test_moe_ffn_tc_matches_fp32_chain or test_moe_combine_tc_matches_fp32"
run dispatch tests/test_gpu_tc.py "test_moe_dispatch_vs_oracle"
run smallconv tests/test_gpu_ops.py "small_conv or stem"
