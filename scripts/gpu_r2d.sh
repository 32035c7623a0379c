#!/bin/bash
mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu "$@" 2> gpurun_out/tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],2), 'passes/step', d['config']['svd_passes_per_step'], 'fused frac', round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; grep "exact rows" gpurun_out/tmp.log | tail -2; }
echo "== 1M exact rows on (trace)"; SCGBM_TRACE=1 run
echo "== 1M exact rows off"; SCGBM_NO_EXACT_ROWS=1 run
echo "== 100k exact rows on (trace)"; SCGBM_TRACE=1 run --cells 100000
echo "== 100k exact rows off"; SCGBM_NO_EXACT_ROWS=1 run --cells 100000
