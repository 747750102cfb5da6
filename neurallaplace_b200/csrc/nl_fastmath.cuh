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
#include "nl_pack2.cuh"

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


// ---- packed (two at a time) versions: same formulas, FMA-pipe work as FADD2 / FMUL2 / FFMA2 (nl_pack2.cuh) ----
// tanh of two hidden units.  The epilogues are bound by the MUFU pipe (16 lanes / clk / SM), so the two share ONE
// reciprocal: 1/(1+E0) = R (1+E1), 1/(1+E1) = R (1+E0) with R = 1/((1+E0)(1+E1)) -- 3 MUFU operations for two tanh
// instead of 4.  E is clamped at 2^60 (tanh is exactly +-1 in float long before) so that the product stays finite.
__device__ __forceinline__ p2::f2 tanh_fast2(p2::f2 z) {
  using namespace p2;
  const f2 E = map2(mul2(z, bc(kTwoLog2e)), [](float v) { return ex2a(fminf(v, 60.f)); });
  const f2 P = add2(bc(1.f), E);
  float p0, p1;
  unpk(P, p0, p1);
  const float R = rcpa(p0 * p1);
  return fma2(bc(-2.f), pk(R * p1, R * p0), bc(1.f));
}

// two head pairs (k, k+1) of the sphere head: oa = (o_a[k], o_a[k+1]), ob = (o_b[k], o_b[k+1])
struct Head2 {
  p2::f2 fre, fim, sn, cs, r, da, db;
};
template <bool NEED_GRAD>
__device__ __forceinline__ Head2 head_forward2(p2::f2 oa, p2::f2 ob) {
  using namespace p2;
  Head2 h;
  const f2 one = bc(1.f), c2 = bc(kTwoLog2e);
  // E clamped at 2^30: 1 -+ tanh is below the 6e-8 floor used further down from there on, and the product of the
  // four (1+E) of the two pairs stays finite -> ONE reciprocal serves all four tanh (MUFU-bound epilogue)
  const f2 Ea = map2(mul2(oa, c2), [](float v) { return ex2a(fminf(v, 30.f)); });
  const f2 Eb = map2(mul2(ob, c2), [](float v) { return ex2a(fminf(v, 30.f)); });
  const f2 pa = add2(one, Ea), pb = add2(one, Eb);
  const f2 pp = mul2(pa, pb);
  float pp0, pp1;
  unpk(pp, pp0, pp1);
  const float R4 = rcpa(pp0 * pp1);
  const f2 R = pk(R4 * pp1, R4 * pp0);  // 1 / (pa pb) of each pair
  const f2 Da = mul2(R, pb), Db = mul2(R, pa);
  // theta = pi * tanh = pi - 2 pi D_a
  const f2 theta = fma2(bc(-2.f * (float)kPi), Da, bc((float)kPi));
  h.sn = map2(theta, [](float v) { return sina(v); });
  h.cs = map2(theta, [](float v) { return cosa(v); });
  const f2 d2 = add2(Db, Db);
  const f2 a1 = map2(d2, [](float v) { return fmaxf(v, 6e-8f); });
  const f2 b1 = map2(mul2(d2, Eb), [](float v) { return fmaxf(v, 6e-8f); });
  if (NEED_GRAD) {
    h.da = mul2(mul2(bc((float)(4.0 * kPi)), Ea), mul2(Da, Da));
    h.db = mul2(bc((float)(0.5 * kPi)), mul2(a1, b1));
  }
  const f2 w = mul2(bc((float)(0.25 * kPi)), zip2(a1, b1, [](float x, float y) { return fminf(x, y); }));
  const f2 w2 = mul2(w, w);
  f2 sw = fma2(w2, bc(2.7557319e-6f), bc(-1.9841270e-4f));
  sw = fma2(w2, sw, bc(8.3333333e-3f));
  sw = fma2(w2, sw, bc(-1.6666667e-1f));
  sw = fma2(mul2(w2, w), sw, w);
  f2 cw = fma2(w2, bc(2.4801587e-5f), bc(-1.3888889e-3f));
  cw = fma2(w2, cw, bc(4.1666667e-2f));
  cw = fma2(w2, cw, bc(-0.5f));
  cw = fma2(w2, cw, one);
  float a1x, a1y, b1x, b1y, swx, swy, cwx, cwy;
  unpk(a1, a1x, a1y);
  unpk(b1, b1x, b1y);
  unpk(sw, swx, swy);
  unpk(cw, cwx, cwy);
  const bool ux = a1x < b1x, uy = a1y < b1y;  // tau_b > 0: r = cot(w) > 1
  const f2 num = pk(ux ? cwx : swx, uy ? cwy : swy);
  const float d0 = ux ? swx : cwx, d1 = uy ? swy : cwy;  // in [sin(pi/4 * 6e-8), 1]: the product cannot over/underflow
  const float Rd = rcpa(d0 * d1);
  h.r = mul2(num, pk(Rd * d1, Rd * d0));
  h.fre = mul2(h.r, h.cs);
  h.fim = mul2(h.r, h.sn);
  return h;
}

// dL/dF -> dL/d(o_a), dL/d(o_b) for the two pairs
__device__ __forceinline__ void head_backward2(const Head2& h, p2::f2 gre, p2::f2 gim, p2::f2* goa, p2::f2* gob) {
  using namespace p2;
  const f2 gtheta = mul2(h.r, sub2(mul2(gim, h.cs), mul2(gre, h.sn)));
  const f2 gr = fma2(gre, h.cs, mul2(gim, h.sn));
  *goa = mul2(gtheta, h.da);
  *gob = mul2(mul2(gr, fma2(h.r, h.r, bc(1.f))), mul2(bc(0.5f), h.db));
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
