#!/usr/bin/env bash
# ncu --set full capture of the fused forward AND the fused backward of one bench step (one ncu call);
# report -> gpurun_out/prof_$1.ncu-rep     usage: scripts/gpu_ncu_both.sh <tag> [bench args]
set -uo pipefail
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/ncu_plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"fused_tc|dehoog_qd" -s ${SKIP:-6} -c ${COUNT:-2} -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_f_${TAG}.log
