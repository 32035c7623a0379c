#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -v -m gpu --timeout=600 -p no:cacheprovider -k "ranks_match or one_caller" 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -80 > gpurun_out/r2s_multi_pytest_$N.log
grep -E "PASSED|FAILED|passed|failed|objective vs oracle" gpurun_out/r2s_multi_pytest_$N.log | cut -c1-400 | tail -25
bash scripts/gpu_r2r.sh $N c5
