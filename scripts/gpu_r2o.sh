#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -q -m gpu -x -k "projection" --timeout=600 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -40
SCGBM_TRACE=1 timeout 900 python scripts/gpu_proj_tc.py 2>&1 | grep -v "tsvd\|iter " | tail -60
