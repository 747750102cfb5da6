// Shared device helpers for the reconstruction hot path: real/complex math wrappers, ILT
// query points s_k(t), Riemann-sphere maps and the per-time-point reduction coefficients.
// Math follows SURVEY.md App. A (reference: torchlaplace/inverse_laplace.py,
// torchlaplace/transformations.py); nothing here is translated from reference source.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nl_b200.h"

#define NL_HD __host__ __device__ __forceinline__

namespace nl {

constexpr double kPi = 3.14159265358979323846;

template <typename R>
struct Cx {
  R re, im;
};

template <typename R>
NL_HD Cx<R> cmul(Cx<R> a, Cx<R> b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename R>
NL_HD Cx<R> cadd(Cx<R> a, Cx<R> b) {
  return {a.re + b.re, a.im + b.im};
}
template <typename R>
NL_HD Cx<R> csub(Cx<R> a, Cx<R> b) {
  return {a.re - b.re, a.im - b.im};
}
template <typename R>
NL_HD Cx<R> cneg(Cx<R> a) {
  return {-a.re, -a.im};
}
template <typename R>
NL_HD Cx<R> cscale(Cx<R> a, R s) {
  return {a.re * s, a.im * s};
}
// Smith's algorithm: robust against overflow of |b|^2 (DeHoog tables span many decades).
template <typename R>
NL_HD Cx<R> cdiv(Cx<R> a, Cx<R> b) {
  // branch-free (selects, two real divisions): with p the larger and s the smaller component of b,
  //   |b.re| >= |b.im|: ((a.re + a.im rat) den, (a.im - a.re rat) den),  rat = b.im/b.re, den = 1/(b.re + b.im rat)
  //   else            : ((a.re rat + a.im) den, (a.im rat - a.re) den),  rat = b.re/b.im, den = 1/(b.re rat + b.im)
  const bool big = fabs(b.re) >= fabs(b.im);
  const R p = big ? b.re : b.im, s = big ? b.im : b.re;
  const R u = big ? a.re : a.im, v = big ? a.im : a.re;
  const R rat = s / p;
  const R den = R(1) / (p + s * rat);
  const R x = v - u * rat;
  return {(u + v * rat) * den, (big ? x : -x) * den};
}
template <typename R>
NL_HD Cx<R> crecip(Cx<R> b) {
  return cdiv(Cx<R>{R(1), R(0)}, b);
}

// ---- scalar math wrappers (accurate libdevice versions; the library is built WITHOUT
// --use_fast_math so float results stay within a few ulp of the reference's ATen kernels) ----
NL_HD float m_tanh(float x) { return tanhf(x); }
NL_HD double m_tanh(double x) { return tanh(x); }
NL_HD float m_tan(float x) { return tanf(x); }
NL_HD double m_tan(double x) { return tan(x); }
NL_HD float m_exp(float x) { return expf(x); }
NL_HD double m_exp(double x) { return exp(x); }
NL_HD float m_asin(float x) { return asinf(x); }
NL_HD double m_asin(double x) { return asin(x); }
NL_HD float m_atan2(float y, float x) { return atan2f(y, x); }
NL_HD double m_atan2(double y, double x) { return atan2(y, x); }
NL_HD float m_sqrt(float x) { return sqrtf(x); }
NL_HD double m_sqrt(double x) { return sqrt(x); }
NL_HD float m_hypot(float x, float y) { return hypotf(x, y); }
NL_HD double m_hypot(double x, double y) { return hypot(x, y); }
NL_HD void m_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
NL_HD void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
NL_HD bool m_isinf(float x) { return isinf(x); }
NL_HD bool m_isinf(double x) { return isinf(x); }

// principal complex square root
template <typename R>
NL_HD Cx<R> csqrt(Cx<R> z) {
  R m = m_hypot(z.re, z.im);
  if (m == R(0)) return {R(0), R(0)};
  R a = m_sqrt((m + fabs(z.re)) * R(0.5));
  R b = z.im / (R(2) * a);
  if (z.re >= R(0)) return {a, b};
  return {fabs(b), z.im >= R(0) ? a : -a};
}

// ---- ILT parameters as the kernels see them -------------------------------------------
template <typename R>
struct IltParams {
  int family;       // NL_FOURIER / NL_AW / NL_DEHOOG
  int n;            // s points per time point
  int head;         // NL_HEAD_SPHERE / NL_HEAD_RAW
  R alpha, log_tol, scale;
  const R* beta;  // [n] complex interleaved (NL_AW)
  const R* eta;   // [n] complex interleaved (NL_AW)
};

// Per-time-point scalars shared by all k.
template <typename R>
struct TimeCtx {
  R t, T;      // T: Fourier/DeHoog contour scale (scale*t unless given), AW: divisor (t unless given)
  R gamma;     // Fourier/DeHoog abscissa  alpha - log(tol)/(scale*T)
  R inv;       // 1/T (Fourier/DeHoog) or 1/t (AW s-points)
  R ratio;     // t/T
  R pref;      // Fourier/DeHoog: exp(gamma*t)/T ; AW: 1/T
};

template <typename R>
NL_HD TimeCtx<R> make_time_ctx(const IltParams<R>& P, R t, const R* Tper, int64_t idx) {
  TimeCtx<R> c;
  c.t = t;
  if (P.family == NL_AW) {
    c.T = Tper ? Tper[idx] : t;
    c.inv = R(1) / t;
    c.gamma = R(0);
    c.ratio = R(1);
    c.pref = R(1) / c.T;
  } else {
    // inverse_laplace.py:1134-1147: T = scale*t; gamma = alpha - log(tol)/(scale*T)
    c.T = Tper ? Tper[idx] : t * P.scale;
    c.gamma = P.alpha - P.log_tol / (P.scale * c.T);
    c.inv = R(1) / c.T;
    c.ratio = t / c.T;
    c.pref = c.inv * m_exp(c.gamma * t);
  }
  return c;
}

// s_k(t): inverse_laplace.py:1146-1147 (Fourier/DeHoog), 1642 (CME/Euler/Gaver), 342-344 (FixedTablot),
// 496 (Stehfest).
template <typename R>
NL_HD Cx<R> s_point(const IltParams<R>& P, const TimeCtx<R>& c, int k) {
  if (P.family == NL_AW) return {c.inv * P.beta[2 * k], c.inv * P.beta[2 * k + 1]};
  return {c.gamma, c.inv * (R(kPi) * R(k))};
}

// MLP input features of one s point: (theta_s, phi_s) on the Riemann sphere
// (transformations.py:56-89) or (Re s, Im s) when use_sphere_projection=False (core.py:131-135).
template <typename R>
NL_HD void s_features(const IltParams<R>& P, Cx<R> s, R* f0, R* f1) {
  if (P.head == NL_HEAD_RAW) {
    *f0 = s.re;
    *f1 = s.im;
    return;
  }
  R a2 = s.im * s.im + s.re * s.re;
  R ratio = m_isinf(a2) ? R(1) : (a2 - R(1)) / (a2 + R(1));
  *f0 = m_atan2(s.im, s.re);
  *f1 = m_asin(ratio);
}

// Reduction coefficient of term k:  x = pref * sum_k (F_re*cr - F_im*ci).
// Fourier (inverse_laplace.py:1198-1208): cr+i*ci = 1/2 (k=0) or exp(i*(t/T)*(k*pi)); AW: eta_k.
template <typename R>
NL_HD Cx<R> reduce_coef(const IltParams<R>& P, const TimeCtx<R>& c, int k) {
  if (P.family == NL_AW) return {P.eta[2 * k], P.eta[2 * k + 1]};
  if (k == 0) return {R(0.5), R(0)};
  R sn, cs;
  m_sincos(c.ratio * (R(k) * R(kPi)), &sn, &cs);
  return {cs, sn};
}

// Head of the canonical laplace_rep_func + spherical_to_complex
// (tests/test_laplace_rep_func.py:56-64, transformations.py:50-53):
//   theta = pi*tanh(o_a), phi = (pi/2)*tanh(o_b), r = tan(phi/2 + pi/4), F = r*(cos theta + i sin theta).
template <typename R>
struct HeadOut {
  R fre, fim;          // F
  R ta, tb, r, sn, cs; // saved for the backward chain rule
};

template <typename R>
NL_HD HeadOut<R> head_forward(int head, R oa, R ob) {
  HeadOut<R> h;
  if (head == NL_HEAD_RAW) {
    h.fre = oa;
    h.fim = ob;
    h.ta = h.tb = h.r = h.sn = h.cs = R(0);
    return h;
  }
  h.ta = m_tanh(oa);
  h.tb = m_tanh(ob);
  R theta = h.ta * R(kPi);
  R phi = h.tb * R(kPi) / R(2.0);
  h.r = m_tan(phi / R(2) + R(kPi) / R(4));
  m_sincos(theta, &h.sn, &h.cs);
  h.fre = h.r * h.cs;
  h.fim = h.r * h.sn;
  return h;
}

// Given dL/dF (gre, gim) return dL/d(o_a), dL/d(o_b)   (SURVEY App. B).
template <typename R>
NL_HD void head_backward(int head, const HeadOut<R>& h, R gre, R gim, R* goa, R* gob) {
  if (head == NL_HEAD_RAW) {
    *goa = gre;
    *gob = gim;
    return;
  }
  R gtheta = h.r * (gim * h.cs - gre * h.sn);
  R gr = gre * h.cs + gim * h.sn;
  R gphi = gr * (R(1) + h.r * h.r) * R(0.5);
  *goa = gtheta * R(kPi) * (R(1) - h.ta * h.ta);
  *gob = gphi * R(kPi * 0.5) * (R(1) - h.tb * h.tb);
}

}  // namespace nl
