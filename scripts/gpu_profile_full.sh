#!/bin/bash
# ncu --set full capture of the steady-state tcgen05 kernels only (launch list: gpu_profile.sh)
# usage: bash scripts/gpu_profile_full.sh <tag> <skip> <count>
TAG=${1:-r01}; SKIP=${2:-262}; CNT=${3:-12}
mkdir -p gpurun_out
CMD="python bench.py --cells 100000 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:fused_tc_kernel|rowsum_tc_kernel|prod_tc_kernel" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
