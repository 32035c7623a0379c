#!/bin/bash
# runs the gpu-marked tests on the GPU box and stores the log
mkdir -p gpurun_out
python -m pytest tests -q -m gpu "$@" 2>&1 | tail -80 | tee gpurun_out/pytest_gpu.log
