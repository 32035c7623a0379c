#!/bin/bash
mkdir -p gpurun_out
for J in 1000000 125000; do
echo "== setup profile J=$J"
SCGBM_TRACE=1 timeout 600 python scripts/gpu_setup_profile.py $J 20 2>&1 | grep -v "iter " | tail -25
done
