#!/usr/bin/env bash
# quick check after a kernel change: the tensor-core parity tests, then the headline bench (no CPU baseline)
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_fused.py tests/test_gpu_parity.py -q -m gpu -x --timeout=600 ${PYTEST_EXTRA:-} 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_q.json'))
    print("value %.1f Mpts/s  ms/step %.3f  fwd %.3f bwd %.3f  e2e %.1f  clocks %s" % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6, d['clocks']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_q.err').read()[-1500:])
PY
done
