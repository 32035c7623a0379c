timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
timeout 300 python scripts/gpu_setup_profile.py 1000000 20 2>&1 | tail -2
run() { timeout 300 python bench.py --cells 113664 --steps 8 --warmup 4 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'fused', round(d['roofline']['ms_per_launch'],3), 'rowsum', round(k['rowsum'],3))"; }
run rider; SCGBM_FUSED_PROD=0 run plain
