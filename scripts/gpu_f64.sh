#!/usr/bin/env bash
# float64 path: parity tests that touch it, then the float64 bench lines of the sweep
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_time_max.py tests/test_gpu_engine.py -q -m gpu -x --timeout=600 ${PYTEST_EXTRA:-} 2>&1 | tail -4
for args in "--alg fourier --dtype f64 --batch 1024" "--alg dehoog --dtype f64 --batch 512" "--alg fourier --dtype f64 --batch 128 --time 100"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-peaks --sustain-s 0 --terms 33 --dobs 1 $args > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_f64.json'))
    print("$args: %.1f Mpts/s  ms/step %.3f  fwd %.3f bwd %.3f" % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms']))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_f64.err').read()[-1500:])
PY
done
