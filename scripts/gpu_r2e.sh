#!/bin/bash
mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu "$@" 2> gpurun_out/tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],2), 'fused ms', round(d['roofline']['ms_per_launch'],3), 'frac', round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items() if v > 0}, d['clocks']['sm_mhz'])"; }
echo "== 1M packed"; run
echo "== 1M scalar"; SCGBM_FUSED_PACKED=0 run
echo "== 1M packed, no rider"; SCGBM_FUSED_PROD=0 run
echo "== 1M scalar, no rider"; SCGBM_FUSED_PACKED=0 SCGBM_FUSED_PROD=0 run
echo "== 100k packed"; run --cells 100000
echo "== 100k scalar"; SCGBM_FUSED_PACKED=0 run --cells 100000
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=1200 2>&1 | tail -40 > gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
