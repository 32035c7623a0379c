#!/bin/bash
# ncu evidence for the bench command (100k-cell shard keeps the replays short).
# usage: bash scripts/gpu_profile.sh <tag> [kernel-regex] [skip] [count]
TAG=${1:-r01}
KRE=${2:-'fused_tc_kernel|rowsum_tc_kernel'}
SKIP=${3:-2}
CNT=${4:-4}
mkdir -p gpurun_out
CMD="python bench.py --cells 100000 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -8
