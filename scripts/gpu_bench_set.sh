#!/usr/bin/env bash
# headline bench + the reference demos' size (eager and graphed host path); lines -> gpurun_out/bench_set.jsonl
set -uo pipefail
mkdir -p gpurun_out; : > gpurun_out/bench_set.jsonl
run() { echo "== $*"; timeout 600 python bench.py "$@" 2> gpurun_out/bench_set.err | tee -a gpurun_out/bench_set.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f Mpts/s  ms/step %.4f  fwd %.4f bwd %.4f | e2e %.1f Mpts/s (%.4f ms) | sustained %s | launches %s | cpu %s' % (d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], (('%.4f ms @ %s MHz' % (d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'))) if d.get('sustained') else None), d['gpu_launches'], d.get('cpu_baseline',{}).get('value')))
print('   roofline', json.dumps({k:v for k,v in d['roofline'].items() if k in ('frac','achieved','precision_mode','whole_step','peaks_measured')}))
" || tail -5 gpurun_out/bench_set.err; }
run --steps 20 --warmup 3
run --steps 50 --warmup 5 --batch 128 --time 100 --no-cpu-baseline --no-peaks --sustain-s 1
run --steps 50 --warmup 5 --batch 128 --time 100 --no-cpu-baseline --no-peaks --sustain-s 1 --engine graph
run --steps 20 --warmup 3 --engine graph --no-cpu-baseline --no-peaks
