#!/bin/bash
# N-GPU box: the multi-GPU parity tests (NCCL ranks, in-process ngpu driver, shim with ngpu), the rider / packed tests,
# then the metric config over all GPUs of the box.  usage: bash scripts/gpu_r2m.sh <ngpu>
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_box_$N.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_r_shim.py -v -m gpu --timeout=600 -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -60 > gpurun_out/r2m_multi_pytest_$N.log
tail -5 gpurun_out/r2m_multi_pytest_$N.log
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_large.py -q -m gpu -k "rider or packed or storage" --timeout=300 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 4 --no-cpu > gpurun_out/r2m_bench_${N}gpu.json 2> gpurun_out/r2m_bench_${N}gpu.log
echo "bench rc=$?"; python - <<P
import json
d=json.loads(open('gpurun_out/r2m_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], d['ms_per_step'], 'ms/step', 'frac', d['roofline']['frac'], 'e2e', d['e2e'] and d['e2e'].get('value'), d['e2e'] and d['e2e'].get('total_seconds'))
print(d.get('kernel_ms_per_step'))
P
