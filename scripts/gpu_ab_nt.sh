#!/usr/bin/env bash
# A/B of an environment knob on the headline bench WITHOUT the parity tests: scripts/gpu_ab_nt.sh VAR v1 v2 ...
set -uo pipefail
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  for i in 1 2; do
    env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-peaks ${BENCH_EXTRA:-} > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_ab.json'))
    print("$VAR=$v value %.1f Mpts/s  ms/step %.3f  fwd %.3f bwd %.3f  e2e %.1f" % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_ab.err').read()[-1500:])
PY
  done
done
