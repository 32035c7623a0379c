#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_multi.py::test_n_ranks_match_one_gpu[$N-kw$2]" "tests/test_gpu_multi.py::test_one_caller_many_gpus[$N]" -v -s -m gpu --timeout=300 -p no:cacheprovider 2>&1 | grep -v "Possible non\|NCCL version\|Maximum number" | grep -E "PASSED|FAILED|SKIPPED|passed|failed|objective vs oracle|angles|Error|differs|mask|truncated|Mismatch|ACTUAL|DESIRED" | cut -c1-600 > gpurun_out/r2w_multi_pytest_$N.log
cat gpurun_out/r2w_multi_pytest_$N.log | tail -30
