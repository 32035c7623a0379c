#!/bin/bash
mkdir -p gpurun_out
short() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d.get('roofline') or {}
print(d['metric'], round(d['value'],4), 'ms/step', round(d['ms_per_step'],3), '| fused', round(r['ms_per_launch'],3), 'ms frac', round(r['frac'],3), 'mix', r.get('mix_peak'), d['clocks']['sm_mhz'])
print('   kernel ms', {k: round(v,2) for k,v in (d.get('kernel_ms_per_step') or {}).items() if v>0}, 'setup', d['config'].get('setup_and_initial_svd_seconds'))
"; }
echo "== c5p"; timeout 900 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
echo "== c5p no rider"; SCGBM_FUSED_PROD=0 timeout 900 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
echo "== c3"; timeout 900 python bench.py --config c3 --steps 10 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
echo "== c5 (M=50)"; timeout 900 python bench.py --config c5 --steps 6 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_r_shim.py -q -m gpu -x --timeout=600 2>&1 | tail -15
