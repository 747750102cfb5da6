#!/usr/bin/env bash
# Builds a profiling variant of the library (per-phase clock64 counters in the 1-stream tcgen05 kernel) into
# /tmp and prints where thread 0 of each CTA spends its cycles.  Not part of the product build.
set -euo pipefail
cd "$(dirname "$0")/.."
# usage: scripts/gpu_phase_profile.sh build   (here, no GPU needed)  |  scripts/gpu_phase_profile.sh   (on the GPU box)
SRC=neurallaplace_b200/csrc
OUT=neurallaplace_b200/lib/libnl_b200_prof.so
if [[ "${1:-}" == "build" ]]; then
  mkdir -p /tmp/nlprof
  FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --extended-lambda -DNL_TC_PROFILE ${NL_EXTRA:-}"
  for f in nl_api nl_elementwise nl_fused_simt nl_fused_tc nl_fused_tc2 nl_tc_selftest nl_dehoog_qd nl_dehoog_fixed nl_optim; do
    nvcc $FLAGS -c $SRC/$f.cu -o /tmp/nlprof/$f.o 2>/dev/null &
  done; wait
  nvcc -shared -o $OUT /tmp/nlprof/*.o -lcudart
  echo "built $OUT"; exit 0
fi
export NL_B200_LIB="$PWD/$OUT"
NL_B200_TC_STREAMS=1 python - <<'PY'
import ctypes, torch, sys
sys.path.insert(0, '.')
from neurallaplace_b200 import laplace_reconstruct, _lib
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
dev = torch.device('cuda', 0)
B, T, terms, D = 4096, 1024, 33, 1
f = LaplaceRepresentationFunc(terms, D, 2).to(dev)
p = torch.randn(B, 2, device=dev, requires_grad=True); t = torch.linspace(20.0/T, 20.0, T, device=dev); g = torch.randn(B, T, D, device=dev)
lib = _lib.load()
for _ in range(2):
    x = laplace_reconstruct(f, p, t, recon_dim=D); x.backward(g)
out = (ctypes.c_longlong * 24)()
lib.nl_debug_tc_profile(out, 1)
x = laplace_reconstruct(f, p, t, recon_dim=D); x.backward(g)
lib.nl_debug_tc_profile(out, 0)
names = ["slot-ctx/prefetch", "1 h1 compute+store+sync (saved bwd: waits dW2,D1)", "2 issue L2 + wait (saved bwd: wait h2 image)", "3 h2 tanh+store+sync", "4 issue L3 + wait", "5 element stage", "6 dO fence+sync+issue", "7 wait dH2", "8 Z2g+aux wait+store+sync", "9 issue dH1/dW2 + wait", "10 Z1g+dp+store+sync", "11 issue D1"]
for k, nm in ((0, "forward"), (12, "backward")):
    v = [out[k+i] for i in range(12)]; tot = sum(v)
    print(nm, "total clocks (sum over CTAs) %.3e  per CTA %.0f" % (tot, tot/296 if k == 0 else tot/148))
    for i in range(12):
        if v[i]: print("   %-28s %5.1f%%" % (names[i], 100.0*v[i]/tot))
tr = (ctypes.c_longlong * 64)()
if hasattr(lib, "nl_debug_tc_trace") and lib.nl_debug_tc_trace(tr) == 0:
    tnames = {}
    for st in (0, 1):
        for sl, nm in enumerate(["tile start", "L3 done + h2 landed", "dO stored+arrive", "dH2 done", "dW3 done (X free)",
                                 "Z2g stored+arrive", "dH1 done + h1 landed", "dW2 done (X free)", "Z1g stored+arrive"]):
            tnames[st * 10 + sl] = "S%d %s" % (st, nm)
    ev = []
    for k in range(3):
        for sl, nm in tnames.items():
            v = tr[sl + 20 * k]
            if v: ev.append((v, "tile+%d %s" % (4 + k, nm)))
    ev.sort()
    if ev:
        t0 = ev[0][0]
        print("event trace of CTA 0, stream 0 (saved backward, two tiles in flight), clocks relative to the first event")
        for v, nm in ev: print("   %8d  %s" % (v - t0, nm))
tc = (ctypes.c_longlong * 256)()
if hasattr(lib, "nl_debug_tc_tileclock") and lib.nl_debug_tc_tileclock(tc) == 0:
    v = [x for x in tc if x]
    if len(v) > 2:
        d = [v[i + 1] - v[i] for i in range(len(v) - 1)]
        print("stream 0 of CTA 0: %d tiles, total %d clk, clocks between consecutive dO stores:" % (len(v), v[-1] - v[0]))
        print("   " + " ".join(str(x) for x in d))
PY
