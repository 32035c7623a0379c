#!/bin/bash
# round 2, first GPU session: the whole gpu test suite on one GPU (new parity helpers, golden fixtures, underflow guard)
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv; } > gpurun_out/box.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=1200 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
