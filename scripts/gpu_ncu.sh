#!/usr/bin/env bash
# ncu capture of the fused kernels (one GPU). Usage: gpu_ncu.sh <tag> [extra bench args]
set -uo pipefail
TAG=${1:-r1}; shift || true
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/ncu_plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_l_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 6 -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_f_${TAG}.log
ls -la gpurun_out | tail -8
