#!/bin/bash
# round 2: compute-sanitizer over the tests of the code added this round (batch entry, row blocks, measurement aids, routed
# vote changes, dense encoder at every tile shape up to 1024) and racecheck over the kernels whose shared-memory use changed
mkdir -p gpurun_out
timeout -k 10 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_batch.py tests/test_gpu_rowsplit.py tests/test_gpu_configs.py tests/test_gpu_fused.py -m gpu -x -q -k "not overflow_flush and not 4096 and not nccl and not dense_1024" > gpurun_out/r2_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_memcheck.log
tail -6 gpurun_out/r2_memcheck.log
timeout -k 10 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_fused.py tests/test_gpu_batch.py -m gpu -x -q -k "vote_fast_matches_oracle or routed or one_call_pipeline_equals_staged or encode_sample_matches_oracle or (batch_records and 3-2)" > gpurun_out/r2_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_racecheck.log
tail -6 gpurun_out/r2_racecheck.log
