#!/bin/bash
# N-GPU box: named configs over N ranks.  usage: bash scripts/gpu_r2r.sh <ngpu> <what...>   what: tests c4 c4small c5 c5p
N=${1:-2}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ "$N" = "1" ] && TR="python"
short() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("no JSON line:", e); sys.exit(0)
r=d.get('roofline') or {}; e=d.get('e2e') or {}
print(d['metric'], round(d['value'],4), d['unit'], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), '| roofline', r.get('achieved') and round(r['achieved'],2), r.get('unit'), 'frac', r.get('frac') and round(r['frac'],3))
print('   e2e', {k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items() if k not in ('checks','call','objective_first_last')})
print('   kernel ms', {k: round(v,2) for k,v in (d.get('kernel_ms_per_step') or r.get('kernel_ms') or {}).items() if v>0}, 'clocks', d.get('clocks',{}).get('sm_mhz'), d.get('clocks',{}).get('reasons'))
print('   config', {k:v for k,v in d['config'].items() if k in ('svd_passes_per_step','setup_and_initial_svd_seconds','genes_kept','irls_sweeps_mean','glm_failures','y_device_format')}, r.get('glm_method'), r.get('tc'))
P
}
for w in "$@"; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_r_shim.py -v -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -60 > gpurun_out/r2r_multi_pytest_$N.log
      grep -E "PASSED|FAILED|SKIPPED|passed|failed" gpurun_out/r2r_multi_pytest_$N.log | tail -25;;
    c4small)
      timeout 900 $TR bench.py --gpus $N --config c4 --cells 200000 --steps 1 --warmup 0 > gpurun_out/r2r_c4small_${N}gpu.json 2> gpurun_out/r2r_c4small_${N}gpu.log; echo "c4small rc=$?"; tail -3 gpurun_out/r2r_c4small_${N}gpu.log | cut -c1-300; short gpurun_out/r2r_c4small_${N}gpu.json;;
    c4)
      timeout 1200 $TR bench.py --gpus $N --config c4 --steps 1 --warmup 0 > gpurun_out/r2r_c4_${N}gpu.json 2> gpurun_out/r2r_c4_${N}gpu.log; echo "c4 rc=$?"; tail -3 gpurun_out/r2r_c4_${N}gpu.log | cut -c1-300; short gpurun_out/r2r_c4_${N}gpu.json;;
    c5)
      timeout 1200 $TR bench.py --gpus $N --config c5 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2r_c5_${N}gpu.json 2> gpurun_out/r2r_c5_${N}gpu.log; echo "c5 rc=$?"; tail -2 gpurun_out/r2r_c5_${N}gpu.log | cut -c1-300; short gpurun_out/r2r_c5_${N}gpu.json;;
    c5p)
      timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 4 --no-cpu > gpurun_out/r2r_c5p_${N}gpu.json 2> gpurun_out/r2r_c5p_${N}gpu.log; echo "c5p rc=$?"; tail -2 gpurun_out/r2r_c5p_${N}gpu.log | cut -c1-300; short gpurun_out/r2r_c5p_${N}gpu.json;;
  esac
done
