#!/bin/bash
# Whole GPU suite WITHOUT -x (every failure is listed), junit + short tracebacks into gpurun_out/; then the driver's own command.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_suite.sh r02a'
tag=${1:-run}
mkdir -p gpurun_out
timeout 2100 python -m pytest tests -q -m gpu -rfEX --tb=short -p no:cacheprovider --durations=25 \
    --junitxml=gpurun_out/gpu_suite_$tag.xml > gpurun_out/gpu_suite_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_suite_$tag.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/gpu_suite_$tag.log | tail -60
