#!/usr/bin/env bash
# Builds a variant of the library with extra nvcc flags into neurallaplace_b200/lib/libnl_b200_$1.so (A/B runs on one
# GPU box: NL_B200_LIB=$PWD/neurallaplace_b200/lib/libnl_b200_$1.so python bench.py ...).   usage: build_variant.sh TAG flags...
set -euo pipefail
cd "$(dirname "$0")/.."
TAG=$1; shift
SRC=neurallaplace_b200/csrc
OUT=neurallaplace_b200/lib/libnl_b200_$TAG.so
mkdir -p /tmp/nlvar_$TAG
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --extended-lambda $*"
for f in nl_api nl_elementwise nl_fused_simt nl_fused_tc nl_fused_tc2 nl_tc_selftest nl_dehoog_qd nl_dehoog_fixed nl_optim; do
  nvcc $FLAGS -c $SRC/$f.cu -o /tmp/nlvar_$TAG/$f.o 2>/dev/null &
done; wait
nvcc -shared -o $OUT /tmp/nlvar_$TAG/*.o -lcudart
echo "built $OUT"
