#!/bin/bash
# ncu --set full capture of late fused-pass launches (with the product rider when it is active)
# usage: bash scripts/gpu_profile_fused.sh <tag> <skip> <count>
TAG=${1:-r01}; SKIP=${2:-5}; CNT=${3:-2}
mkdir -p gpurun_out
CMD="python bench.py --cells 100000 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain3_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:fused_tc_kernel" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
