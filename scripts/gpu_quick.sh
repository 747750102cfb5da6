#!/usr/bin/env bash
# quick perf check: streams setting(s) given as args, e.g. "0 2"
set -uo pipefail
mkdir -p gpurun_out
for S in "$@"; do
  NL_B200_TC_STREAMS=$S timeout 300 python -m pytest tests/test_gpu_tc_fused.py -q -m gpu -x --timeout=120 2>&1 | tail -1
  NL_B200_TC_STREAMS=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q_$S.json 2> gpurun_out/bench_q_$S.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_q_$S.json'))
    print("streams=$S value %.1f Mpts/s  ms/step %.3f  fwd %.3f bwd %.3f  e2e %.1f  clocks %s" % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6, d['clocks']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_q_$S.err').read()[-1500:])
PY
done
