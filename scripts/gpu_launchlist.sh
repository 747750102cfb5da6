#!/usr/bin/env bash
# ncu launch list (device time per kernel) of two bench steps -> stdout + gpurun_out/launches_$1.csv
TAG=${1:-ll}; shift || true
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peaks --sustain-s 0 "$@" > gpurun_out/plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peaks --sustain-s 0 "$@" > gpurun_out/ncu_${TAG}.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_${TAG}.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]
for r in rows[hi+1:]:
    d=dict(zip(h,r))
    print("%-72s %10s %s" % (d["Kernel Name"][:72], d["Metric Value"], d["Metric Unit"]))
PY
