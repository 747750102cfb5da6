#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
for S in 1 2 0; do
  NL_B200_TC_STREAMS=$S timeout 600 python -m pytest tests/test_gpu_tc_fused.py tests/test_gpu_tc.py -q -m gpu -x --timeout=120 > gpurun_out/pytest_tc_$S.log 2>&1; echo "streams=$S pytest tc rc=$?"
  grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_tc_$S.log | head -5
  NL_B200_TC_STREAMS=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $* > gpurun_out/bench_tc_$S.json 2> gpurun_out/bench_tc_$S.err; echo "bench rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_tc_$S.json'))
    print("streams=$S value %.1f Mpts/s  ms/step %.3f  fwd %.3f bwd %.3f  e2e %.1f Mpts/s  frac %.4f" % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6, d['roofline']['frac']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_tc_$S.err').read()[-2000:])
PY
done
