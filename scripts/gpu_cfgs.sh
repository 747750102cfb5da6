#!/usr/bin/env bash
# short bench of a few configurations given as quoted arg strings, e.g. gpu_cfgs.sh "--dobs 1" "--terms 65"
for a in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $a | python -c '
import json,sys
d=json.loads(sys.stdin.read()); c=d["config"]
print(c["workload"].split(" ")[0], "%.1f Mpts/s %.3f ms fwd %.3f bwd %.3f e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], c["fwd_ms"], c["bwd_ms"], d["e2e"]["value"]/1e6))'
done
