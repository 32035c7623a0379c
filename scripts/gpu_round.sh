#!/bin/bash
# one GPU-box session: box facts, gpu tests, bench at 100k and 1M cells
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv; } > gpurun_out/box.txt 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cells 100000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.log
cat gpurun_out/bench_100k.json | head -c 600; echo
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.log
echo "rc=$?"; tail -5 gpurun_out/bench_1m.log; cat gpurun_out/bench_1m.json | head -c 1500; echo
