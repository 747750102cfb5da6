// Packed FP32 pairs (PTX add / sub / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2, sm_100+).
// One FFMA2 occupies the FMA pipe for two cycles -- the same FLOP rate as two FFMA -- but takes ONE issue slot
// (scripts/mb/ffma2_mb.cu, measured on B200: 8 FFMA 8.2 clk, 8 FFMA2 16.1 clk per warp and SMSP, and they overlap
// with MUFU work).  The fused kernels' epilogues are bound by issue slots and the MUFU pipe with the FMA pipe a
// third busy, so evaluating two heads / two hidden units per instruction removes ~a third of their instructions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nl {
namespace p2 {

struct f2 {
  unsigned long long v;
};

__device__ __forceinline__ f2 pk(float x, float y) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ void unpk(f2 a, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
}
__device__ __forceinline__ float lo(f2 a) {
  float x, y;
  unpk(a, x, y);
  return x;
}
__device__ __forceinline__ float hi(f2 a) {
  float x, y;
  unpk(a, x, y);
  return y;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
// element-wise helpers for the operations that have no packed form (MUFU, min / max, select)
template <typename F>
__device__ __forceinline__ f2 map2(f2 a, F fn) {
  float x, y;
  unpk(a, x, y);
  return pk(fn(x), fn(y));
}
template <typename F>
__device__ __forceinline__ f2 zip2(f2 a, f2 b, F fn) {
  float ax, ay, bx, by;
  unpk(a, ax, ay);
  unpk(b, bx, by);
  return pk(fn(ax, bx), fn(ay, by));
}

}  // namespace p2
}  // namespace nl
