// De Hoog / Knight / Stokes accelerated Fourier inversion: quotient-difference table ->
// continued-fraction coefficients d_0..d_2M -> Pade recurrence with the improved remainder.
// One thread runs the whole recurrence for one (batch, time, d_obs) element, so the algorithm
// stays parallel over batch and time (BASELINE.json north_star).  The recursion order is the
// one of SURVEY.md App. A.6 (reference: torchlaplace/inverse_laplace.py:785-866); it is
// ill-conditioned, so the order matters for parity.
//
// All maps a_k -> w = A*/B* are holomorphic, so reverse mode propagates plain complex
// derivatives dw/d(.) and only the last step takes x = pref * Re(w):
//   dL/dRe(a_k) = g*pref*Re(dw/da_k),  dL/dIm(a_k) = -g*pref*Im(dw/da_k).
//
// Tables live in a caller-provided scratch area, element `idx` of slot `slot` at
// scr[idx*stride + slot] (coalesced across the threads of a warp).
#pragma once
#include "nl_common.cuh"

namespace nl {

template <typename R>
struct Scratch {
  Cx<R>* base;
  int64_t stride;
  int64_t slot;
  NL_HD Cx<R>& operator[](int idx) const { return base[(int64_t)idx * stride + slot]; }
  // the same slot, `off` entries further on
  NL_HD Scratch<R> at(int off) const { return Scratch<R>{base + (int64_t)off * stride, stride, slot}; }
};

// number of complex scratch entries per slot
__host__ __device__ inline int dehoog_scratch_entries(int n, bool backward) {
  int M = (n - 1) / 2;
  int LD = 2 * M + 2;
  return backward ? (2 * M + 1 + 8) * LD : 3 * LD;
}

// The recurrence comes in parts so that the fixed-contour variant (DeHoog.fixed_line_integrate,
// inverse_laplace.py:587-701: ONE quotient-difference table per contour, the Pade recurrence per time point)
// runs exactly the same arithmetic as the per-element kernels:
//   dehoog_qd_coeffs   : a_k -> d_0..d_2M, two in-place columns                  (forward only)
//   dehoog_pade        : d, z -> w = A*/B*                                          (forward only)
//   dehoog_qd_store    : a_k -> every q^r / e^r column + d                          (kept for the reverse sweep)
//   dehoog_pade_fwd_bwd: d, z -> w and dw/dd_i                                      (A_i, B_i history kept)
//   dehoog_qd_reverse  : dw/dd_i -> dw/da_k through the stored columns

// a[k]: value k of the Laplace-domain samples F_k.  Writes d_0..d_2M to scr[2 LD + i]; scr needs 3 LD entries.
template <typename R, typename AFn>
__host__ __device__ void dehoog_qd_coeffs(int M, AFn a, Scratch<R> scr) {
  const int LD = 2 * M + 2;
  const int Q = 0, E = LD, Dd = 2 * LD;  // q[2M], e[2M+1], d[2M+1]
  const Cx<R> half_a0 = cscale(a(0), R(0.5));
  scr[Dd + 0] = half_a0;
  for (int i = 0; i < 2 * M; ++i) scr[Q + i] = (i == 0) ? cdiv(a(1), half_a0) : cdiv(a(i + 1), a(i));
  for (int i = 0; i <= 2 * M; ++i) scr[E + i] = {R(0), R(0)};
  for (int r = 1; r <= M; ++r) {
    const int mr = 2 * (M - r) + 1;
    scr[Dd + 2 * r - 1] = cneg(scr[Q + 0]);
    for (int i = 0; i < mr; ++i) scr[E + i] = cadd(csub(scr[Q + i + 1], scr[Q + i]), scr[E + i + 1]);
    scr[E + mr] = {R(0), R(0)};
    scr[Dd + 2 * r] = cneg(scr[E + 0]);
    if (r < M) {
      for (int i = 0; i < mr; ++i) scr[Q + i] = cdiv(cmul(scr[Q + i + 1], scr[E + i + 1]), scr[E + i]);
    }
  }
}

// Pade recurrence with the improved remainder for one z; d(i) returns d_i.
template <typename R, typename DFn>
__host__ __device__ Cx<R> dehoog_pade(int M, Cx<R> z, DFn d) {
  Cx<R> Am1 = {R(0), R(0)}, A = d(0), Bm1 = {R(1), R(0)}, B = {R(1), R(0)};
  for (int i = 1; i < 2 * M; ++i) {
    // reference order: (d_i * A_{i-2}) * z
    const Cx<R> di = d(i);
    Cx<R> An = cadd(A, cmul(cmul(di, Am1), z));
    Cx<R> Bn = cadd(B, cmul(cmul(di, Bm1), z));
    Am1 = A; A = An; Bm1 = B; B = Bn;
  }
  const Cx<R> one = {R(1), R(0)};
  Cx<R> d2m1 = d(2 * M - 1), d2m = d(2 * M);
  Cx<R> h = cscale(cadd(one, cmul(csub(d2m1, d2m), z)), R(0.5));
  Cx<R> U = cadd(one, cdiv(cmul(d2m, z), cmul(h, h)));
  Cx<R> rem = cmul(cneg(h), csub(one, csqrt(U)));
  Cx<R> As = cadd(A, cmul(rem, Am1));
  Cx<R> Bs = cadd(B, cmul(rem, Bm1));
  return cdiv(As, Bs);
}

// Forward only (two in-place columns).  Returns w = A*/B*.
template <typename R, typename AFn>
__host__ __device__ Cx<R> dehoog_forward(int M, Cx<R> z, AFn a, Scratch<R> scr) {
  dehoog_qd_coeffs<R>(M, a, scr);
  const Scratch<R> dd = scr.at(2 * (2 * M + 2));
  return dehoog_pade<R>(M, z, [&](int i) { return dd[i]; });
}

// Quotient-difference table with every column kept.  tab layout (units of LD = 2M+2):
//   q^r r=0..M-1 | e^r r=0..M | d        ((2M+2) LD entries)
template <typename R, typename AFn>
__host__ __device__ void dehoog_qd_store(int M, AFn a, Scratch<R> tab) {
  const int LD = 2 * M + 2;
  const int Q = 0, E = M * LD, Dd = (2 * M + 1) * LD;
  const Cx<R> zero = {R(0), R(0)};
  const Cx<R> half_a0 = cscale(a(0), R(0.5));
  tab[Dd + 0] = half_a0;
  for (int i = 0; i < 2 * M; ++i) tab[Q + i] = (i == 0) ? cdiv(a(1), half_a0) : cdiv(a(i + 1), a(i));
  for (int i = 0; i < LD; ++i) tab[E + i] = zero;  // e^0 == 0
  for (int r = 1; r <= M; ++r) {
    const int mr = 2 * (M - r) + 1;
    const int qp = Q + (r - 1) * LD, ep = E + (r - 1) * LD, ec = E + r * LD, qc = Q + r * LD;
    for (int i = 0; i < mr; ++i) tab[ec + i] = cadd(csub(tab[qp + i + 1], tab[qp + i]), tab[ep + i + 1]);
    tab[ec + mr] = zero;
    tab[Dd + 2 * r - 1] = cneg(tab[qp + 0]);
    tab[Dd + 2 * r] = cneg(tab[ec + 0]);
    if (r < M) {
      for (int i = 0; i < mr; ++i) tab[qc + i] = cdiv(cmul(tab[qp + i + 1], tab[ec + i + 1]), tab[ec + i]);
    }
  }
}

// Pade recurrence for one z with the A_i / B_i history kept (hist: A at [0, LD), B at [LD, 2 LD)), then its
// reverse sweep: dbar[i] = dw/dd_i, i = 0..2M (OVERWRITTEN).  d(i) returns d_i.  Returns w.
template <typename R, typename DFn>
__host__ __device__ Cx<R> dehoog_pade_fwd_bwd(int M, Cx<R> z, DFn d, Scratch<R> dbar, Scratch<R> hist) {
  const int LD = 2 * M + 2;
  const int AA = 0, BB = LD;
  const Cx<R> zero = {R(0), R(0)}, one = {R(1), R(0)};
  Cx<R> Am1 = zero, A = d(0), Bm1 = one, B = one;
  hist[AA + 0] = A;
  hist[BB + 0] = B;
  for (int i = 1; i < 2 * M; ++i) {
    const Cx<R> di = d(i);
    Cx<R> An = cadd(A, cmul(cmul(di, Am1), z));
    Cx<R> Bn = cadd(B, cmul(cmul(di, Bm1), z));
    Am1 = A; A = An; Bm1 = B; B = Bn;
    hist[AA + i] = A;
    hist[BB + i] = B;
  }
  Cx<R> d2m1 = d(2 * M - 1), d2m = d(2 * M);
  Cx<R> h = cscale(cadd(one, cmul(csub(d2m1, d2m), z)), R(0.5));
  Cx<R> h2 = cmul(h, h);
  Cx<R> U = cadd(one, cdiv(cmul(d2m, z), h2));
  Cx<R> S = csqrt(U);
  Cx<R> rem = cmul(cneg(h), csub(one, S));
  Cx<R> As = cadd(A, cmul(rem, Am1));
  Cx<R> Bs = cadd(B, cmul(rem, Bm1));
  Cx<R> w = cdiv(As, Bs);

  // ---- reverse: Pade part
  for (int i = 0; i <= 2 * M; ++i) dbar[i] = zero;
  Cx<R> Asb = crecip(Bs);
  Cx<R> Bsb = cneg(cdiv(w, Bs));
  // A* = A_{2M-1} + rem*A_{2M-2}
  Cx<R> ab1 = Asb, ab0 = cmul(Asb, rem);  // adjoints of A_{2M-1} (complete) and A_{2M-2} (partial)
  Cx<R> bb1 = Bsb, bb0 = cmul(Bsb, rem);
  Cx<R> remb = cadd(cmul(Asb, Am1), cmul(Bsb, Bm1));
  // rem = -h*(1-S), S = sqrt(U), U = 1 + d2m*z/h^2
  Cx<R> hb = cmul(remb, cneg(csub(one, S)));
  Cx<R> Sb = cmul(remb, h);
  Cx<R> Ub = cdiv(Sb, cscale(S, R(2)));
  Cx<R> zh2 = cdiv(z, h2);
  Cx<R> d2mb = cmul(Ub, zh2);
  hb = csub(hb, cmul(Ub, cdiv(cscale(cmul(d2m, z), R(2)), cmul(h2, h))));
  // h = (1 + (d_{2M-1} - d_{2M}) z)/2
  Cx<R> hz = cscale(cmul(hb, z), R(0.5));
  dbar[2 * M - 1] = hz;
  dbar[2 * M] = csub(d2mb, hz);
  for (int i = 2 * M - 1; i >= 1; --i) {
    // A_i = A_{i-1} + d_i z A_{i-2}
    Cx<R> Ai2 = (i >= 2) ? hist[AA + i - 2] : zero;
    Cx<R> Bi2 = (i >= 2) ? hist[BB + i - 2] : one;
    Cx<R> di = d(i);
    Cx<R> g = cmul(cadd(cmul(ab1, Ai2), cmul(bb1, Bi2)), z);
    dbar[i] = cadd(dbar[i], g);
    Cx<R> diz = cmul(di, z);
    Cx<R> na = cmul(ab1, diz), nb = cmul(bb1, diz);
    ab0 = cadd(ab0, ab1);
    bb0 = cadd(bb0, bb1);
    ab1 = ab0; ab0 = na;
    bb1 = bb0; bb0 = nb;
  }
  dbar[0] = cadd(dbar[0], ab1);  // A_0 = d_0 ; B_0 = 1 is a constant
  return w;
}

// Reverse sweep through the stored quotient-difference table (tab: dehoog_qd_store layout): dbar[i] = dw/dd_i
// in, emit_grad(k, dw/da_k) out, k = 0..2M.  adj: 4 LD entries of ping-pong adjoint columns.
template <typename R, typename AFn, typename GFn>
__host__ __device__ void dehoog_qd_reverse(int M, AFn a, GFn emit_grad, Scratch<R> tab, Scratch<R> dbar,
                                           Scratch<R> adj) {
  const int LD = 2 * M + 2;
  const int Q = 0, E = M * LD;
  const int QB0 = 0, QB1 = LD, EB0 = 2 * LD, EB1 = 3 * LD;
  const Cx<R> zero = {R(0), R(0)};
  const Cx<R> half_a0 = cscale(a(0), R(0.5));
  int qb_cur = QB0, qb_prev = QB1, eb_cur = EB0, eb_prev = EB1;
  for (int i = 0; i < LD; ++i) {
    adj[qb_cur + i] = zero;  // adjoint of q^M (does not exist) -- unused
    adj[eb_cur + i] = zero;
  }
  adj[eb_cur + 0] = cneg(dbar[2 * M]);  // d_{2M} = -e^M_0
  for (int r = M; r >= 1; --r) {
    const int mr = 2 * (M - r) + 1;
    const int qp = Q + (r - 1) * LD, ec = E + r * LD, qc = Q + r * LD;
    for (int i = 0; i < LD; ++i) {
      adj[qb_prev + i] = zero;
      adj[eb_prev + i] = zero;
    }
    adj[qb_prev + 0] = cneg(dbar[2 * r - 1]);               // d_{2r-1} = -q^{r-1}_0
    if (r > 1) adj[eb_prev + 0] = cneg(dbar[2 * (r - 1)]);  // d_{2(r-1)} = -e^{r-1}_0
    if (r < M) {
      // q^r_i = q^{r-1}_{i+1} * e^r_{i+1} / e^r_i
      for (int i = 0; i < mr; ++i) {
        Cx<R> g = adj[qb_cur + i];
        Cx<R> ei = tab[ec + i], ei1 = tab[ec + i + 1], qpi1 = tab[qp + i + 1];
        Cx<R> gi = cdiv(g, ei);
        adj[qb_prev + i + 1] = cadd(adj[qb_prev + i + 1], cmul(gi, ei1));
        if (i + 1 < mr) adj[eb_cur + i + 1] = cadd(adj[eb_cur + i + 1], cmul(gi, qpi1));
        adj[eb_cur + i] = csub(adj[eb_cur + i], cmul(gi, tab[qc + i]));
      }
    }
    // e^r_i = q^{r-1}_{i+1} - q^{r-1}_i + e^{r-1}_{i+1}
    for (int i = 0; i < mr; ++i) {
      Cx<R> g = adj[eb_cur + i];
      adj[qb_prev + i + 1] = cadd(adj[qb_prev + i + 1], g);
      adj[qb_prev + i] = csub(adj[qb_prev + i], g);
      if (r > 1) adj[eb_prev + i + 1] = cadd(adj[eb_prev + i + 1], g);
    }
    int tq = qb_cur; qb_cur = qb_prev; qb_prev = tq;
    int te = eb_cur; eb_cur = eb_prev; eb_prev = te;
  }
  // ---- q^0_i = a_{i+1}/a_i (i>=1), q^0_0 = a_1/(a_0/2) ; d_0 = a_0/2
  Cx<R> carry = cscale(dbar[0], R(0.5));  // running adjoint of a_i, starts with a_0
  for (int i = 0; i < 2 * M; ++i) {
    Cx<R> g = adj[qb_cur + i];
    Cx<R> den = (i == 0) ? half_a0 : a(i);
    Cx<R> gi = cdiv(g, den);
    Cx<R> back = cmul(gi, tab[Q + i]);  // d q/d den = -q/den
    if (i == 0) back = cscale(back, R(0.5));
    emit_grad(i, csub(carry, back));
    carry = gi;
  }
  emit_grad(2 * M, carry);
}

// Forward with the full table kept, then the reverse sweep.  On return emit_grad(k, .) has been called
// once per k with dw/da_k (holomorphic derivative); returns w.
template <typename R, typename AFn, typename GFn>
__host__ __device__ Cx<R> dehoog_forward_backward(int M, Cx<R> z, AFn a, GFn emit_grad, Scratch<R> scr) {
  const int LD = 2 * M + 2;
  // layout (in units of LD): q^r r=0..M-1 | e^r r=0..M | d | dbar | A | B | qbA | qbB | ebA | ebB
  const int Dd = (2 * M + 1) * LD, Db = Dd + LD, AA = Db + LD, QB0 = AA + 2 * LD;
  dehoog_qd_store<R>(M, a, scr);
  const Scratch<R> dd = scr.at(Dd);
  const Cx<R> w = dehoog_pade_fwd_bwd<R>(M, z, [&](int i) { return dd[i]; }, scr.at(Db), scr.at(AA));
  dehoog_qd_reverse<R>(M, a, emit_grad, scr, scr.at(Db), scr.at(QB0));
  return w;
}

}  // namespace nl
