#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels with warp-level rings / staging
mkdir -p gpurun_out
timeout -k 10 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_fused.py tests/test_gpu_point_encoder.py -m gpu -x -q -k "vote_fast_matches_oracle or rot_hist or knn_selects or one_call_pipeline_equals_staged or encode_sample_matches_oracle" > gpurun_out/racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck.log
tail -40 gpurun_out/racecheck.log
