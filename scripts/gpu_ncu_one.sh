#!/usr/bin/env bash
# ncu --set full capture of ONE kernel (regex $1) of a short bench run; report -> gpurun_out/prof_$2.ncu-rep
set -uo pipefail
PAT=$1; TAG=$2; shift 2
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/ncu_plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${PAT} -s 3 -c 1 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_f_${TAG}.log
