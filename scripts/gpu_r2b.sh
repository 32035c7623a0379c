#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu_tall_diag2.py 2>&1 | grep -v Warn | tee gpurun_out/tall_diag2.log | tail -20
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_large.py tests/test_gpu_golden.py -q -m gpu --timeout=900 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
tail -30 gpurun_out/r2b_pytest.log
