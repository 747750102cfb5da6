// De Hoog recurrence with compile-time M (terms = 2M+1) and every working column in registers.
// Same arithmetic, in the same order, as nl_dehoog.cuh (reference: torchlaplace/inverse_laplace.py:785-866) --
// tests/test_host_math.py checks the two against each other on the host -- but laid out for the GPU:
//   * forward: the two quotient-difference columns q, e live in registers (loops over the column are unrolled, so the
//     array indices are static; the loop over columns stays rolled -- fully unrolled code ran at the speed of
//     instruction-cache misses), the Pade recurrence consumes d_{2r-1}, d_{2r} as each column produces them;
//     nothing is spilled to memory at all;
//   * forward + reverse: the forward additionally streams the table (q^r, e^r) to `tab` (global memory, one
//     coalesced 8-byte access per thread and entry) and the Pade history A_i, B_i to `hist` (shared memory);
//     the reverse sweep keeps the two adjoint columns in registers and updates them IN PLACE, column by column,
//     interleaved with the reverse Pade steps that produce the adjoints of d_{2r-1}, d_{2r-2} just in time.
// Per element the only memory traffic is the table: ~4.2 KB written and ~6.5 KB read for 33 terms, against the
// ~45 KB of scratch traffic of the generic-M version.
#pragma once
#include "nl_common.cuh"

namespace nl {

template <int M>
struct DehoogRegLayout {
  static constexpr int LD = 2 * M + 2;
  static constexpr int kTabEntries = (2 * M + 1) * LD;  // q^r at r*LD + i (r < M), e^r at (M + r)*LD + i (1 <= r <= M)
  static constexpr int kHB = 2 * M - 2;                 // Pade history kept for i <= 2M-3 (the reverse reads A_{i-2})
  static constexpr int kHistEntries = 2 * kHB;          // A_i at i, B_i at kHB + i
  NL_HD static constexpr int q(int r, int i) { return r * LD + i; }
  NL_HD static constexpr int e(int r, int i) { return (M + r) * LD + i; }
};

// Forward only: returns w = A*/B*.
template <typename R, int M, typename AFn>
NL_HD Cx<R> dehoog_reg_forward(Cx<R> z, AFn a) {
  const Cx<R> zero = {R(0), R(0)}, one = {R(1), R(0)};
  Cx<R> q[2 * M], e[2 * M + 1];
  const Cx<R> half_a0 = cscale(a(0), R(0.5));
  {
    Cx<R> ai = half_a0;
#pragma unroll
    for (int i = 0; i < 2 * M; ++i) {
      const Cx<R> an = a(i + 1);
      q[i] = cdiv(an, ai);
      ai = an;
    }
  }
#pragma unroll
  for (int i = 0; i <= 2 * M; ++i) e[i] = zero;
  Cx<R> Am1 = zero, A = half_a0, Bm1 = one, B = one;
  Cx<R> d2m1 = zero, d2m = zero;
#pragma unroll 1
  for (int r = 1; r <= M; ++r) {
    const int mr = 2 * (M - r) + 1;  // (uniform across the warp: the guards below are uniform branches)
    const Cx<R> dodd = cneg(q[0]);
#pragma unroll
    for (int i = 0; i < 2 * M - 1; ++i)
      if (i < mr) e[i] = cadd(csub(q[i + 1], q[i]), e[i + 1]);
#pragma unroll
    for (int i = 1; i <= 2 * M - 1; ++i)
      if (i == mr) e[i] = zero;
    const Cx<R> deven = cneg(e[0]);
    if (r < M) {
#pragma unroll
      for (int i = 0; i < 2 * M - 1; ++i)
        if (i < mr) q[i] = cdiv(cmul(q[i + 1], e[i + 1]), e[i]);
    }
    {  // Pade step i = 2r-1
      const Cx<R> An = cadd(A, cmul(cmul(dodd, Am1), z));
      const Cx<R> Bn = cadd(B, cmul(cmul(dodd, Bm1), z));
      Am1 = A; A = An; Bm1 = B; B = Bn;
    }
    if (r < M) {  // Pade step i = 2r
      const Cx<R> An = cadd(A, cmul(cmul(deven, Am1), z));
      const Cx<R> Bn = cadd(B, cmul(cmul(deven, Bm1), z));
      Am1 = A; A = An; Bm1 = B; B = Bn;
    } else {
      d2m1 = dodd;
      d2m = deven;
    }
  }
  const Cx<R> h = cscale(cadd(one, cmul(csub(d2m1, d2m), z)), R(0.5));
  const Cx<R> U = cadd(one, cdiv(cmul(d2m, z), cmul(h, h)));
  const Cx<R> rem = cmul(cneg(h), csub(one, csqrt(U)));
  const Cx<R> As = cadd(A, cmul(rem, Am1));
  const Cx<R> Bs = cadd(B, cmul(rem, Bm1));
  return cdiv(As, Bs);
}

// Forward keeping the table, then the reverse sweep.  emit_grad(k, dw/da_k) is called once per k; returns w.
// tab(idx) / hist(idx) return references to this element's slot of entry idx.
// stage: a per-element ring of kStageEntries table entries filled asynchronously (cp.async on the GPU) --
// issue(k, idx) starts the copy of table entry idx into ring entry k, commit() closes a group, wait<N>() waits
// until at most N groups are pending, get(k) reads ring entry k.
constexpr int kDhChunk = 4;                            // table rows of one column fetched per group
constexpr int kDhDepth = 4;                            // groups in flight
constexpr int kDhStageEntries = kDhDepth * 3 * kDhChunk;
template <typename R, int M, typename AFn, typename GFn, typename Tab, typename Hist, typename Stage>
NL_HD Cx<R> dehoog_reg_forward_backward(Cx<R> z, AFn a, GFn emit_grad, Tab tab, Hist hist, Stage stage) {
  using Ly = DehoogRegLayout<M>;
  const Cx<R> zero = {R(0), R(0)}, one = {R(1), R(0)};
  const Cx<R> half_a0 = cscale(a(0), R(0.5));
  Cx<R> Am1 = zero, A = half_a0, Bm1 = one, B = one;
  Cx<R> d2m1 = zero, d2m = zero;
  {
    Cx<R> q[2 * M], e[2 * M + 1];
    {
      Cx<R> ai = half_a0;
#pragma unroll
      for (int i = 0; i < 2 * M; ++i) {
        const Cx<R> an = a(i + 1);
        q[i] = cdiv(an, ai);
        tab(Ly::q(0, i)) = q[i];
        ai = an;
      }
    }
#pragma unroll
    for (int i = 0; i <= 2 * M; ++i) e[i] = zero;
    if (Ly::kHB > 0) hist(0) = A;
    if (Ly::kHB > 0) hist(Ly::kHB + 0) = B;
#pragma unroll 1
    for (int r = 1; r <= M; ++r) {
      const int mr = 2 * (M - r) + 1;
      const Cx<R> dodd = cneg(q[0]);
#pragma unroll
      for (int i = 0; i < 2 * M - 1; ++i) {
        if (i < mr) {
          e[i] = cadd(csub(q[i + 1], q[i]), e[i + 1]);
          tab(Ly::e(r, i)) = e[i];
        }
      }
#pragma unroll
      for (int i = 1; i <= 2 * M - 1; ++i)
        if (i == mr) e[i] = zero;
      const Cx<R> deven = cneg(e[0]);
      if (r < M) {
#pragma unroll
        for (int i = 0; i < 2 * M - 1; ++i) {
          if (i < mr) {
            q[i] = cdiv(cmul(q[i + 1], e[i + 1]), e[i]);
            tab(Ly::q(r, i)) = q[i];
          }
        }
      }
      {
        const Cx<R> An = cadd(A, cmul(cmul(dodd, Am1), z));
        const Cx<R> Bn = cadd(B, cmul(cmul(dodd, Bm1), z));
        Am1 = A; A = An; Bm1 = B; B = Bn;
        if (2 * r - 1 <= 2 * M - 3) {  // (the reverse sweep reads A_{i-2}, i <= 2M-1)
          hist(2 * r - 1) = A;
          hist(Ly::kHB + 2 * r - 1) = B;
        }
      }
      if (r < M) {
        const Cx<R> An = cadd(A, cmul(cmul(deven, Am1), z));
        const Cx<R> Bn = cadd(B, cmul(cmul(deven, Bm1), z));
        Am1 = A; A = An; Bm1 = B; B = Bn;
        if (2 * r <= 2 * M - 3) {
          hist(2 * r) = A;
          hist(Ly::kHB + 2 * r) = B;
        }
      } else {
        d2m1 = dodd;
        d2m = deven;
      }
    }
  }
  const Cx<R> h = cscale(cadd(one, cmul(csub(d2m1, d2m), z)), R(0.5));
  const Cx<R> h2 = cmul(h, h);
  const Cx<R> U = cadd(one, cdiv(cmul(d2m, z), h2));
  const Cx<R> S = csqrt(U);
  const Cx<R> rem = cmul(cneg(h), csub(one, S));
  const Cx<R> As = cadd(A, cmul(rem, Am1));
  const Cx<R> Bs = cadd(B, cmul(rem, Bm1));
  const Cx<R> w = cdiv(As, Bs);

  // ---- reverse: remainder
  const Cx<R> Asb = crecip(Bs);
  const Cx<R> Bsb = cneg(cdiv(w, Bs));
  Cx<R> ab1 = Asb, ab0 = cmul(Asb, rem);  // adjoints of A_{2M-1} (complete) and A_{2M-2} (partial)
  Cx<R> bb1 = Bsb, bb0 = cmul(Bsb, rem);
  const Cx<R> remb = cadd(cmul(Asb, Am1), cmul(Bsb, Bm1));
  Cx<R> hb = cmul(remb, cneg(csub(one, S)));
  const Cx<R> Sb = cmul(remb, h);
  const Cx<R> Ub = cdiv(Sb, cscale(S, R(2)));
  const Cx<R> zh2 = cdiv(z, h2);
  const Cx<R> d2mb = cmul(Ub, zh2);
  hb = csub(hb, cmul(Ub, cdiv(cscale(cmul(d2m, z), R(2)), cmul(h2, h))));
  const Cx<R> hz = cscale(cmul(hb, z), R(0.5));
  const Cx<R> dbar_2m = csub(d2mb, hz);

  // one reverse Pade step: A_i = A_{i-1} + d_i z A_{i-2}; returns the adjoint of d_i
  auto pade_rev = [&](int i, Cx<R> di) {
    const Cx<R> Ai2 = (i >= 2) ? hist(i - 2) : zero;
    const Cx<R> Bi2 = (i >= 2) ? hist(Ly::kHB + i - 2) : one;
    Cx<R> g = cmul(cadd(cmul(ab1, Ai2), cmul(bb1, Bi2)), z);
    if (i == 2 * M - 1) g = cadd(hz, g);
    const Cx<R> diz = cmul(di, z);
    const Cx<R> na = cmul(ab1, diz), nb = cmul(bb1, diz);
    ab0 = cadd(ab0, ab1);
    bb0 = cadd(bb0, bb1);
    ab1 = ab0; ab0 = na;
    bb1 = bb0; bb0 = nb;
    return g;
  };

  // ---- reverse: quotient-difference table, columns r = M .. 1; qb / eb = adjoints of the current q / e column.
  // The table entries a column needs (e^r_j, q^r_j, q^{r-1}_j) are fetched kDhChunk rows at a time through the
  // asynchronous stage, kDhDepth groups ahead of the arithmetic and across column boundaries, so the sweep never
  // waits for a global load with only two CTAs of four warps on an SM (registers: 2 x 66 for the adjoint columns).
  constexpr int kChunk = kDhChunk, kDepth = kDhDepth, kPer = 3 * kDhChunk;
  constexpr int kLast = 2 * M - 1;  // largest row index of any column
  int ir = M - 1, ij = 0, iss = 0, cons = 0;
  auto issue_next = [&]() {
    if (ir >= 1) {
      const int sl = (iss % kDepth) * kPer;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        int j = ij + u;
        if (j > kLast) j = kLast;  // (beyond the column: fetched, never used)
        stage.issue(sl + 3 * u + 0, Ly::e(ir, j));
        stage.issue(sl + 3 * u + 1, Ly::q(ir, j));
        stage.issue(sl + 3 * u + 2, Ly::q(ir - 1, j));
      }
      ij += kChunk;
      if (ij > 2 * (M - ir) + 1) {
        ij = 0;
        --ir;
      }
    }
    stage.commit();
    ++iss;
  };
#pragma unroll 1
  for (int k = 0; k < kDepth; ++k) issue_next();
  Cx<R> qb[2 * M + 1], eb[2 * M + 1];
#pragma unroll
  for (int i = 0; i <= 2 * M; ++i) qb[i] = eb[i] = zero;
  // d_{2r-1} = -q^{r-1}_0, d_{2r-2} = -e^{r-1}_0: loaded one column ahead
  Cx<R> dq_cur = tab(Ly::q(M - 1, 0)), de_cur = (M > 1) ? tab(Ly::e(M - 1, 0)) : zero;
#pragma unroll 1
  for (int r = M; r >= 1; --r) {
    const int mr = 2 * (M - r) + 1;
    Cx<R> dq_nxt = zero, de_nxt = zero;
    if (r > 1) {
      dq_nxt = tab(Ly::q(r - 2, 0));
      if (r > 2) de_nxt = tab(Ly::e(r - 2, 0));
    }
    const Cx<R> dbar_a = pade_rev(2 * r - 1, cneg(dq_cur));
    Cx<R> dbar_b = zero;
    if (r > 1) dbar_b = pade_rev(2 * r - 2, cneg(de_cur));
    dq_cur = dq_nxt;
    de_cur = de_nxt;
    if (r == M) {
      const Cx<R> g = cneg(dbar_2m);  // d_{2M} = -e^M_0
      qb[1] = g;
      qb[0] = csub(cneg(dbar_a), g);
      if (r > 1) {
        eb[1] = g;
        eb[0] = cneg(dbar_b);
      }
    } else {
      // q^r_i = q^{r-1}_{i+1} e^r_{i+1} / e^r_i ;  e^r_i = q^{r-1}_{i+1} - q^{r-1}_i + e^{r-1}_{i+1}
      Cx<R> gi_prev = zero, ebf_prev = zero;
#pragma unroll
      for (int j0 = 0; j0 <= kLast; j0 += kChunk) {
        if (j0 > mr) break;
        stage.template wait<kDepth - 1>();
        const int sl = (cons % kDepth) * kPer;
        Cx<R> ev[kChunk], qcv[kChunk], qpv[kChunk];
#pragma unroll
        for (int u = 0; u < kChunk; ++u) {
          ev[u] = stage.get(sl + 3 * u + 0);
          qcv[u] = stage.get(sl + 3 * u + 1);
          qpv[u] = stage.get(sl + 3 * u + 2);
        }
#pragma unroll
        for (int u = 0; u < kChunk; ++u) {
          const int j = j0 + u;
          if (j <= kLast) {
            const bool lt = j < mr, le = j <= mr;
            Cx<R> gi_j = zero, ebf_j = zero, e_j = zero;
            if (lt) {
              e_j = ev[u];
              gi_j = cdiv(qb[j], e_j);
              Cx<R> tt = eb[j];
              if (j >= 1) tt = cadd(tt, cmul(gi_prev, qpv[u]));
              ebf_j = csub(tt, cmul(gi_j, qcv[u]));
            }
            if (le) {
              Cx<R> nq;
              if (j == 0) nq = csub(cneg(dbar_a), ebf_j);
              else if (lt) nq = csub(cadd(cmul(gi_prev, e_j), ebf_prev), ebf_j);
              else nq = cadd(cmul(gi_prev, zero), ebf_prev);
              qb[j] = nq;
              if (r > 1) eb[j] = (j == 0) ? cneg(dbar_b) : ebf_prev;
            }
            gi_prev = gi_j;
            ebf_prev = ebf_j;
          }
        }
        ++cons;
        issue_next();
      }
    }
  }
  stage.template wait<0>();
  // ---- q^0_i = a_{i+1}/a_i (i >= 1), q^0_0 = a_1/(a_0/2) ; d_0 = a_0/2 (its adjoint is the last ab1)
  Cx<R> carry = cscale(ab1, R(0.5));
#pragma unroll
  for (int i = 0; i < 2 * M; ++i) {
    const Cx<R> den = (i == 0) ? half_a0 : a(i);
    const Cx<R> gi = cdiv(qb[i], den);
    Cx<R> back = cmul(gi, tab(Ly::q(0, i)));  // d q / d den = -q / den
    if (i == 0) back = cscale(back, R(0.5));
    emit_grad(i, csub(carry, back));
    carry = gi;
  }
  emit_grad(2 * M, carry);
  return w;
}

}  // namespace nl
