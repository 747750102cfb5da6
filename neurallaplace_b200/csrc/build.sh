#!/usr/bin/env bash
# Build libnl_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --extended-lambda
       -Xptxas -v)
SRCS=(nl_api nl_elementwise nl_fused_simt nl_fused_tc nl_fused_tc2 nl_tc_selftest nl_dehoog_qd nl_dehoog_fixed nl_optim)
pids=()
for f in "${SRCS[@]}"; do
  stale=0
  if [[ ! -f "$HERE/obj/$f.o" || "$HERE/$f.cu" -nt "$HERE/obj/$f.o" ]]; then stale=1; fi
  for h in "$HERE"/*.cuh "$HERE"/*.h "$HERE"/../../include/*.h; do
    [[ -f "$HERE/obj/$f.o" && "$h" -nt "$HERE/obj/$f.o" ]] && stale=1
  done
  if [[ $stale == 1 ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" > "$HERE/obj/$f.log" 2>&1 \
        || { cat "$HERE/obj/$f.log"; rm -f "$HERE/obj/$f.o"; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || rc=1; }; done
[[ $rc == 0 ]] || { echo "build failed"; exit 1; }
objs=(); for f in "${SRCS[@]}"; do objs+=("$HERE/obj/$f.o"); done
"$NVCC" -shared -o "$OUT/libnl_b200.so" "${objs[@]}" -lcudart
echo "built $OUT/libnl_b200.so"
