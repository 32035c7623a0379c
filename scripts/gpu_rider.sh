#!/bin/bash
# product rider of the fused pass: parity tests, check against the plain pass, timing at 100k and 1M cells
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fit.py -q -x -k "rider" 2>&1 | tail -15 > gpurun_out/rider_pytest.log
tail -5 gpurun_out/rider_pytest.log
SCGBM_CHECK_FUSED=1 timeout 300 python bench.py --cells 100000 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/rider_check.json 2> gpurun_out/rider_check.log
echo "rc=$?"; grep "scgbm-check" gpurun_out/rider_check.log | head -8
timeout 300 python bench.py --cells 100000 --steps 8 --warmup 4 --no-e2e --no-cpu > gpurun_out/rider_100k.json 2> gpurun_out/rider_100k.log
echo "rc=$?"; head -c 900 gpurun_out/rider_100k.json; echo
SCGBM_FUSED_PROD=0 timeout 300 python bench.py --cells 100000 --steps 8 --warmup 4 --no-e2e --no-cpu > gpurun_out/norider_100k.json 2> gpurun_out/norider_100k.log
echo "rc=$?"; head -c 900 gpurun_out/norider_100k.json; echo
