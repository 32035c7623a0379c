#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -q -m gpu -k "projection" --timeout=600 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -15
timeout 900 python scripts/gpu_proj_tc.py 2>&1 | grep -v "tsvd\|iter \|Warning\|warnings.warn\|Possible non" | tail -8
for J in 1000000 125000; do
echo "== setup profile J=$J"
SCGBM_TRACE=1 timeout 600 python scripts/gpu_setup_profile.py $J 20 2>&1 | grep -v "iter " | tail -25
done
