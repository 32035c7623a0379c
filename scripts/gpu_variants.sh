#!/bin/bash
# A/B of prebuilt library variants (variants/libscgbm_*.so, not tracked), interleaved repeats:
# fused pass with and without the product rider at 113 664 cells (six full waves of cell tiles)
run() { timeout 300 python bench.py --cells 113664 --steps 8 --warmup 4 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'fused', round(d['roofline']['ms_per_launch'],3), 'rowsum', round(k['rowsum'],3))"; }
cp scgbm_b200/libscgbm.so /tmp/libscgbm_keep.so
for rep in 1 2; do
for v in $(ls variants/*.so); do
  cp $v scgbm_b200/libscgbm.so
  run "$(basename $v) rider"; SCGBM_FUSED_PROD=0 run "$(basename $v) plain"
done
done
cp /tmp/libscgbm_keep.so scgbm_b200/libscgbm.so
