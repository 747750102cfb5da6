#!/usr/bin/env bash
# builds libnl_b200.so variants with -DNL_TC_PASS_MODE=m into /tmp and runs dbg_xerr + quick bench for each
set -uo pipefail
SRC=neurallaplace_b200/csrc
cp neurallaplace_b200/lib/libnl_b200.so /tmp/orig.so
trap 'cp /tmp/orig.so neurallaplace_b200/lib/libnl_b200.so' EXIT
for m in "$@"; do
  mkdir -p /tmp/pm$m
  FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --extended-lambda -DNL_TC_PASS_MODE=$m"
  for f in nl_api nl_elementwise nl_fused_simt nl_fused_tc nl_fused_tc2 nl_tc_selftest nl_dehoog_qd nl_optim; do nvcc $FLAGS -c $SRC/$f.cu -o /tmp/pm$m/$f.o 2>/dev/null & done; wait
  nvcc -shared -o neurallaplace_b200/lib/libnl_b200.so /tmp/pm$m/*.o -lcudart
  echo "=== pass mode $m"; python scripts/dbg_xerr.py | tail -1; bash scripts/gpu_cfgs.sh "--dobs 1"
done
