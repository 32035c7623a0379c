#!/bin/bash
mkdir -p gpurun_out
for tol in 1e-4 3e-5 1e-5; do
  echo "=== SCGBM_SVD_TOL=$tol"
  SCGBM_SVD_TOL=$tol python scripts/gpu_tall_diag.py 2>&1 | grep -v -i warn | grep -A3 "rider=force {}"
  SCGBM_SVD_TOL=$tol timeout 600 python bench.py --cells 100000 --steps 6 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/b100k_$tol.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('100k: ms/step', round(d['ms_per_step'],2), 'passes/step', d['config']['svd_passes_per_step'], 'fused frac', round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
  grep "setup" gpurun_out/b100k_$tol.log
done 2>&1 | tee gpurun_out/r2c.log
for tol in 1e-4 1e-5; do
  echo "=== 1M SCGBM_SVD_TOL=$tol"
  SCGBM_SVD_TOL=$tol timeout 900 python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu 2> gpurun_out/b1m_$tol.log | tee gpurun_out/b1m_$tol.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1M: ms/step', round(d['ms_per_step'],2), 'passes/step', d['config']['svd_passes_per_step'], 'fused frac', round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['clocks'])"
  grep "setup" gpurun_out/b1m_$tol.log
done 2>&1 | tee -a gpurun_out/r2c.log
