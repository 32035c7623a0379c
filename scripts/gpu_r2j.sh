#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_r_shim.py -q -m gpu -x --timeout=300 2>&1 | grep -E "Error|error|assert|stderr|Assertion" | head -20
short() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d.get('roofline') or {}
print(d['metric'], round(d['value'],4), 'ms/step', round(d['ms_per_step'],3), '| fused', round(r['ms_per_launch'],3), 'ms frac', round(r['frac'],3), d['clocks']['sm_mhz'], 'passes/step', d['config']['svd_passes_per_step'])
print('   kernel ms', {k: round(v,2) for k,v in (d.get('kernel_ms_per_step') or {}).items() if v>0}, 'launches', {k: round(v,1) for k,v in d['launches_per_step'].items() if v>0})
"; }
echo "== c5 (M=50) coop jacobi from 64"; SCGBM_JACOBI_COOP_FROM=64 timeout 900 python bench.py --config c5 --steps 6 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
echo "== c5 (M=50) svd_extra 6 (block 56)"; timeout 900 python bench.py --config c5 --steps 6 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
