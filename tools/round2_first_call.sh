#!/bin/bash
# First GPU call of the next round: run the suites that were written after round 1's GPU budget was spent (all xfail(strict=False):
# an XPASS line means the kernels are good and the marker can be removed), then take first numbers for the new model families.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/round2_first_call.log and gpurun_out/models_r2.json.
mkdir -p gpurun_out
{
  for f in tests/test_gpu_zz_predictor.py tests/test_gpu_zz_nms_large.py tests/test_gpu_zz_model_v0_1.py tests/test_gpu_zz_pose.py tests/test_gpu_zz_segment.py tests/test_gpu_zz_obb.py tests/test_gpu_zz_classify.py tests/test_gpu_zz_postproc.py tests/test_gpu_zz_latent.py tests/test_gpu_zz_gated.py; do
    echo "=== $f"
    timeout 600 python -m pytest "$f" -q -rxXf --tb=short -p no:cacheprovider 2>&1 | tail -60
  done
} > gpurun_out/round2_first_call.log 2>&1
timeout 600 python tools/bench_models.py 32 > gpurun_out/models_r2.json 2> gpurun_out/models_r2.err
tail -5 gpurun_out/round2_first_call.log
