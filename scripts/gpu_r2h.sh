#!/bin/bash
# new bench.py smoke at every single-GPU config + the implicit-W gate experiment (no-store rider)
mkdir -p gpurun_out
short() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d.get('roofline') or {}
print(d['metric'], round(d['value'],4), d['unit'], 'ms/step', round(d['ms_per_step'],3), '| roofline', r.get('achieved') and round(r['achieved'],2), r.get('unit'), 'frac', r.get('frac') and round(r['frac'],3), 'mix', r.get('mix_peak'), 'frac_mix', r.get('frac_of_mix_peak') and round(r['frac_of_mix_peak'],3))
e=d.get('e2e') or {}
print('   e2e', {k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items() if k not in ('checks','call','objective_first_last')})
print('   checks', e.get('checks'))
print('   kernel ms', {k: round(v,2) for k,v in (d.get('kernel_ms_per_step') or {}).items() if v>0}, 'cpu', (d.get('cpu_baseline') or {}).get('value'))
"; }
for c in c2 c2se c3 c3se; do echo "== $c"; timeout 900 python bench.py --config $c --steps 5 --warmup 3 2> gpurun_out/bench_$c.log | tee gpurun_out/bench_$c.json | short; tail -2 gpurun_out/bench_$c.log | cut -c1-300; done
echo "== c5p no e2e"; timeout 900 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/bench_c5p_quick.log | tee gpurun_out/bench_c5p_quick.json | short
echo "== c5p gate: rider without the plane stores = implicit-W product R^T Q reading only Y"; SCGBM_DBG_NO_STORE=1 timeout 900 python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/tmp.log | short
echo "== shim"; timeout 600 python -m pytest tests/test_r_shim.py -q -m gpu --timeout=300 2>&1 | tail -5
