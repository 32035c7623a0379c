#!/bin/bash
mkdir -p gpurun_out
for T in 1e-5 3e-6 1e-6; do
echo "== warm tol $T"
SCGBM_SVD_TOL_WARM=$T timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu 2> gpurun_out/tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('ms/step', round(d['ms_per_step'],2), 'e2e fit s', round(e['fit_seconds'],3), 'device ms', round(e['device_ms'],1), 'accepted', e['accepted'], 'iters', e['iterations'], 'obj last', e['objective_first_last'][1], 'unconv', e['checks'].get('svd_unconverged_calls'))
"
done
