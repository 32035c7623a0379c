#!/bin/bash
# product rider of the fused pass: parity tests, check against the plain pass, timing with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_large.py -q -x -k "rider" 2>&1 | tail -15 > gpurun_out/rider_pytest.log
tail -3 gpurun_out/rider_pytest.log
SCGBM_CHECK_FUSED=1 timeout 300 python bench.py --cells 100000 --steps 2 --warmup 3 --no-e2e --no-cpu 2>&1 | grep "scgbm-check" | head -3
run() { timeout 300 python bench.py --cells $2 --steps 8 --warmup 4 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],3), 'fused', round(d['roofline']['ms_per_launch'],3), round(d['roofline']['frac'],3))"; }
for c in 100000 113664 1000000; do run rider $c; SCGBM_FUSED_PROD=0 run plain $c; done
