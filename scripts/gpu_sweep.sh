#!/usr/bin/env bash
# sweep of BASELINE.json configs[4] (terms x d_obs x family) -> gpurun_out/sweep.jsonl
set -uo pipefail
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-peaks --sustain-s 0 "$@" >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err || echo "{\"failed\": \"$*\"}" >> gpurun_out/sweep.jsonl; }
for terms in 33 65 129; do for d in 1 2 4 16; do run --alg fourier --terms $terms --dobs $d; done; done
for alg in cme euler fixed_tablot stehfest dehoog; do run --alg $alg --terms 33 --dobs 1; done
run --alg dehoog --terms 33 --dobs 1 --kernel simt
run --alg dehoog --terms 33 --dobs 2
run --alg dehoog --terms 65 --dobs 1
run --alg fourier --terms 33 --dobs 1 --tmode per_row
run --alg fourier --terms 33 --dobs 1 --dtype f64 --batch 1024
run --alg dehoog --terms 33 --dobs 1 --dtype f64 --batch 512
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'failed' in d: print("FAILED", d['failed']); continue
    c=d['config']
    print(f"{c['workload'].split(' ')[0]:55s} {c['kernel_family'][:8]:8s} {d['value']/1e6:9.1f} Mpts/s  {d['ms_per_step']:9.3f} ms  fwd {c['fwd_ms']:.3f} bwd {c['bwd_ms']:.3f}  tensor-frac {d['roofline']['frac']:.4f}")
PY
