// MUFU / FMA-pipe evaluation of the laplace_rep_func heads for the tensor-core path.
//
// The epilogue (tanh, tanh heads, sphere -> complex, weighted reduction) is the binding unit of the
// fused kernels (~300 transcendentals against ~17 kFLOP of GEMM per point at d_obs=1,
// 33 terms), so it is written for instruction count while staying ~1e-6 accurate:
//   tanh(x)    = 1 - 2/(1 + e^{2x})            ex2.approx + rcp.approx    (abs err ~1e-7)
//   1 -/+ tanh = 2/(1+E), 2E/(1+E)             relatively accurate near the poles of the sphere; the two heads
//                                               of a pair share one rcp: 1/(1+Ea) = (1+Eb) / ((1+Ea)(1+Eb))
//   sin, cos   of theta = pi*tanh in (-pi, pi)  sin/cos.approx            (abs err 2^-21)
//   r = tan(phi/2 + pi/4) = cot(w) or tan(w),  w = (pi/4) * min(1 - tau, 1 + tau) in (0, pi/4]
//                                               degree-9/8 Taylor on the FMA pipe + one rcp.approx:
//                                               RELATIVE accuracy ~2e-7 for every r (a MUFU tan
//                                               would lose it near the pole, where r ~ 1/(1-tau)).
#pragma once
#include "nl_common.cuh"

namespace nl {
namespace fast {

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpa(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sina(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float cosa(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kTwoLog2e = 2.885390081777927f;  // 2 / ln 2

// E = e^{2x}, D = 1/(1+E).  Only the upper clamp is needed: it keeps E finite so that E*D (= 1 - D) never
// evaluates inf*0; for x -> -inf, E = 0 and D = 1 are exact.
__device__ __forceinline__ void exp_pair(float x, float* E, float* Dn) {
  *E = ex2a(fminf(x * kTwoLog2e, 100.f));
  *Dn = rcpa(1.f + *E);
}

// tanh(x) = 1 - 2/(1+e^{2x}); no clamp at all: E = inf gives D = 0 -> 1, E = 0 gives D = 1 -> -1.
__device__ __forceinline__ float tanh_fast(float x) {
  return fmaf(-2.f, rcpa(1.f + ex2a(x * kTwoLog2e)), 1.f);
}

struct Head {
  float fre, fim;  // F = r (cos theta + i sin theta)
  float sn, cs, r;
  float da, db;    // d theta / d o_a = pi (1 - tau_a^2),  d phi / d o_b = (pi/2)(1 - tau_b^2)
};

__device__ __forceinline__ Head head_forward(int head, float oa, float ob) {
  Head h;
  if (head == NL_HEAD_RAW) {
    h.fre = oa; h.fim = ob; h.sn = 0.f; h.cs = 0.f; h.r = 0.f; h.da = 1.f; h.db = 1.f;
    return h;
  }
  // The two tanh heads share ONE reciprocal: D_a = 1/(1+E_a) = R (1+E_b), D_b = R (1+E_a) with
  // R = 1/((1+E_a)(1+E_b)) -- the epilogue is MUFU-bound (7 -> 6 MUFU ops per head pair).  E is clamped at 2^60
  // (tanh is exactly +-1 in float there) so that the product stays finite.
  const float Ea = ex2a(fminf(oa * kTwoLog2e, 60.f)), Eb = ex2a(fminf(ob * kTwoLog2e, 60.f));
  const float pa = 1.f + Ea, pb = 1.f + Eb;
  const float R = rcpa(pa * pb);
  const float Da = R * pb, Db = R * pa;
  const float ta = fmaf(-2.f, Da, 1.f);
  h.da = (float)(4.0 * kPi) * Ea * Da * Da;
  const float theta = (float)kPi * ta;
  h.sn = sina(theta);
  h.cs = cosa(theta);
  // 1 - tau_b and 1 + tau_b, each relatively accurate; clamp = one float ulp of 1 (tanh saturates there)
  const float a1 = fmaxf(2.f * Db, 6e-8f);
  const float b1 = fmaxf(2.f * Eb * Db, 6e-8f);
  h.db = (float)(0.5 * kPi) * a1 * b1;
  const float w = (float)(0.25 * kPi) * fminf(a1, b1);
  const float w2 = w * w;
  float sw = fmaf(w2, 2.7557319e-6f, -1.9841270e-4f);
  sw = fmaf(w2, sw, 8.3333333e-3f);
  sw = fmaf(w2, sw, -1.6666667e-1f);
  sw = fmaf(w2 * w, sw, w);
  float cw = fmaf(w2, 2.4801587e-5f, -1.3888889e-3f);
  cw = fmaf(w2, cw, 4.1666667e-2f);
  cw = fmaf(w2, cw, -0.5f);
  cw = fmaf(w2, cw, 1.f);
  const bool upper = a1 < b1;  // tau_b > 0: r = cot(w) > 1
  h.r = (upper ? cw : sw) * rcpa(upper ? sw : cw);
  h.fre = h.r * h.cs;
  h.fim = h.r * h.sn;
  return h;
}

// dL/dF -> dL/d(o_a), dL/d(o_b)
__device__ __forceinline__ void head_backward(int head, const Head& h, float gre, float gim, float* goa, float* gob) {
  if (head == NL_HEAD_RAW) {
    *goa = gre;
    *gob = gim;
    return;
  }
  const float gtheta = h.r * (gim * h.cs - gre * h.sn);
  const float gr = gre * h.cs + gim * h.sn;
  *goa = gtheta * h.da;
  *gob = gr * fmaf(h.r, h.r, 1.f) * 0.5f * h.db;
}

// atan(t) for t in [0, 1]: t * P(t^2), degree-8 least-squares fit on Chebyshev nodes; max abs error 9e-8 in float.
__device__ __forceinline__ float atan01(float t) {
  const float z = t * t;
  float p = 2.9327620e-3f;
  p = fmaf(p, z, -1.6413191e-2f);
  p = fmaf(p, z, 4.3278243e-2f);
  p = fmaf(p, z, -7.5569004e-2f);
  p = fmaf(p, z, 1.0667487e-1f);
  p = fmaf(p, z, -1.4211106e-1f);
  p = fmaf(p, z, 1.9993694e-1f);
  p = fmaf(p, z, -3.3333138e-1f);
  p = fmaf(p, z, 1.0f);
  return p * t;
}

// Riemann-sphere features of s = re + i im (transformations.py:56-89): theta = atan2(im, re),
// phi = asin((|s|^2 - 1) / (|s|^2 + 1)) = 2 atan(|s|) - pi/2.  Used where the features feed a bf16x3 operand
// image (per-row time grids, point_layer1_kernel): the image keeps 16 mantissa bits, the ~2e-7 absolute error of
// this version is far below that, and it costs ~50 instructions per s point against ~110 for atan2f + asinf
// with their IEEE divisions.  (The asin form loses accuracy as |s| grows -- the float32 reference is ~1e-5 off
// at |s| ~ 500 -- the atan form does not.)
__device__ __forceinline__ void sphere_features(float re, float im, float* theta, float* phi) {
  constexpr float kHalfPi = 1.57079632679489662f, kPiF = 3.14159265358979324f;
  const float a2 = fmaf(im, im, re * re);
  float ri;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"(a2));
  ri = ri * fmaf(-0.5f * a2 * ri, ri, 1.5f);  // one Newton step: 1 / |s|
  const float r = a2 * ri;                    // |s|
  const bool big = r > 1.f;
  const float pr = atan01(big ? ri : r);
  float ph = big ? fmaf(-2.f, pr, kHalfPi) : fmaf(2.f, pr, -kHalfPi);
  if (!(a2 < 3.0e38f)) ph = kHalfPi;   // |s|^2 overflowed: the reference's isinf branch (ratio = 1)
  if (a2 == 0.f) ph = -kHalfPi;
  const float ax = fabsf(re), ay = fabsf(im);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float pt = atan01(mx > 0.f ? __fdividef(mn, mx) : 0.f);
  if (ay > ax) pt = kHalfPi - pt;
  if (re < 0.f) pt = kPiF - pt;
  *theta = copysignf(pt, im);
  *phi = ph;
}

}  // namespace fast
}  // namespace nl
