#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest "tests/test_gpu_multi.py::test_n_ranks_match_one_gpu[2-kw0]" "tests/test_gpu_multi.py::test_one_caller_many_gpus[2]" "tests/test_gpu_multi.py::test_sharded_projection_matches_one_gpu[2]" -v -s -m gpu --timeout=300 -p no:cacheprovider 2>&1 | grep -v "Possible non\|NCCL version\|Maximum number" | grep -E "PASSED|FAILED|SKIPPED|passed|failed|objective vs oracle|angles|Error|differs|mask" | cut -c1-500 > gpurun_out/r2_2gpu_check.log
cat gpurun_out/r2_2gpu_check.log
