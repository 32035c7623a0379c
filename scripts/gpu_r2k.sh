#!/bin/bash
# round-2 state check: whole gpu suite, then the metric config with AUTO (u8) / pinned u16, rider on / off, packed on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -15 > gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
short() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d.get('roofline') or {}
print(d['metric'], round(d['value'],4), 'ms/step', round(d['ms_per_step'],3), '| fused', round(r['ms_per_launch'],3), 'ms frac', round(r['frac'],3), 'GB', round(r['bytes_per_launch']/1e9,1), 'mix', {k: round(v) for k,v in r.get('mix_peak',{}).items()}, d['clocks']['sm_mhz'], d['config']['y_device_format'], 'passes', d['config']['svd_passes_per_step'])
print('   kernel ms', {k: round(v,2) for k,v in (d.get('kernel_ms_per_step') or {}).items() if v>0}, 'launches', {k: round(v,1) for k,v in d['launches_per_step'].items() if v>0})
"; }
run() { echo "== $1"; shift; env "$@" timeout 900 python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short; }
run "c5p AUTO (u8)" A=1
run "c5p u16" SCGBM_BENCH_STORAGE=u16
run "c5p AUTO no rider" SCGBM_FUSED_PROD=0
run "c5p u16 no rider" SCGBM_BENCH_STORAGE=u16 SCGBM_FUSED_PROD=0
run "c5p AUTO rider packed" SCGBM_FUSED_PACKED=1
run "c5p AUTO no rider unpacked" SCGBM_FUSED_PROD=0 SCGBM_FUSED_PACKED=0
