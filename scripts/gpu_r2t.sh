#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout=600 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -15 > gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
bash scripts/gpu_r2r.sh 1 c5p
