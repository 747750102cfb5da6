#!/usr/bin/env bash
# weak + strong scaling at N ranks of one box: scripts/gpu_scale.sh N   (lines -> gpurun_out/scale_N.jsonl)
set -uo pipefail
N=$1
mkdir -p gpurun_out; : > gpurun_out/scale_$N.jsonl
for mode in ${MODES:-weak strong}; do
  if [[ $N == 1 ]]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-peaks --scaling $mode 2> gpurun_out/scale_err.log | tee -a gpurun_out/scale_$N.jsonl > /dev/null
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-peaks --scaling $mode 2> gpurun_out/scale_err.log | grep '^{' | tee -a gpurun_out/scale_$N.jsonl > /dev/null
  fi
done
python - <<PY
import json
for l in open('gpurun_out/scale_$N.jsonl'):
    d=json.loads(l)
    print("N=%d %-6s value %.1f Mpts/s ms/step %.4f (fwd %.3f bwd %.3f) e2e %.1f Mpts/s (%.4f ms) sustained %.4f ms rows/rank %d" % (d['n_gpus'], d['scaling'], d['value']/1e6, d['ms_per_step'], d['config']['fwd_ms'], d['config']['bwd_ms'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['sustained']['ms_per_step'], d['config']['per_gpu_batch']))
PY
tail -3 gpurun_out/scale_err.log
