#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2v_box_$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_r_shim.py -v -s -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|Possible non\|NCCL version" | grep -E "PASSED|FAILED|SKIPPED|passed|failed|objective vs oracle|Error" | cut -c1-500 > gpurun_out/r2v_multi_pytest_$N.log
cat gpurun_out/r2v_multi_pytest_$N.log | tail -40
