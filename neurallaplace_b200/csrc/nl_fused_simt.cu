// Fused laplace_reconstruct forward / backward on CUDA cores (FFMA / DFMA), any dtype, any ILT
// family, shared or per-row time grids.  This is the float64 path (tcgen05 has no fp64 kind) and
// the general-shape path; the float32 canonical shape runs on tensor cores (nl_fused_tc.cu).
//
// One persistent CTA (256 threads) walks tiles of `TT x BB` (time x batch) rows.  Per tile it
//   1. generates the s points of its time slots, projects them on the Riemann sphere and applies
//      the s-part of layer 1 (S1 = W1s . u_t) -- amortised over the BB batch rows of a slot;
//   2. evaluates layers 1..2 of the canonical laplace_rep_func into shared memory;
//   3. for each d_obs evaluates the 2n outputs of layer 3, the tanh heads, the sphere->complex map
//      and the complex-weighted reduction over terms (or the De Hoog recurrence) without the
//      (batch x time x terms x d_obs) complex intermediate ever leaving the SM;
//   4. (backward) recomputes 1-3, and back-propagates: weight gradients go to a CTA-private slab
//      in the workspace (plain read-modify-write by an owning thread -> deterministic), reduced
//      over CTAs by a second tiny kernel; dp uses atomics.
// Rows are mapped to lanes (row = lane + 32*i), output columns to warps, so weight reads are
// warp-uniform (one L1 transaction) and activation reads from shared memory are conflict-free.
//
// Math: SURVEY.md App. A/B; reference core.py:104-150, inverse_laplace.py, transformations.py.
#include <algorithm>
#include <cstdlib>

#include "nl_dehoog.cuh"
#include "nl_kernels.h"

namespace nl {

// float32: 256 threads, CUDA-core FFMA GEMMs with rows <-> lanes.  float64: 512 threads and the two GEMM shapes on the
// tensor pipe (DMMA, mma.sync.m8n8k4.f64): with 32 rows per tile (shared memory) the FFMA-style loop issues 9 loads per
// 8 DFMA and eight warps cannot hide their latency (measured 1.7 TFLOP/s of GEMM work, 5 % of the FP64 peak).
template <typename R>
struct SimtCfg {
  static constexpr int kThreads = sizeof(R) == 8 ? 512 : 256;
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kPad = sizeof(R) == 8 ? 4 : 1;  // row padding of the shared-memory matrices (see smem_layout)
};
constexpr int kKCU = 32;  // s points per feature chunk of layer 1

template <typename R>
struct FusedArgs {
  IltParams<R> P;
  int B, T, K, H, D, n, t_mode;
  int BB, TT, nbt, ntt;
  const R* p;
  const R* t;
  const R *W1, *b1, *W2, *b2, *W3, *b3;
  R* x;
  const R* gx;
  R* slab;
  int64_t slab_len;
  R* dp;
  // De Hoog: the recurrence runs in its own kernels (nl_dehoog_qd.cu); the forward instantiation writes
  // F[d][k][t][b] here, the backward one reads dL/dF (already scaled by grad_x * pref, imaginary part negated)
  Cx<R>* fplane;
  const Cx<R>* dfplane;
};

struct SmemLayout {
  int ldh, ldo, ldu;
  size_t H1, H2, G, S1, O, U, Coef, Ctx, Pm, X, total;  // offsets in units of R
};

// pad = 1 (float32): odd row strides, conflict-free for the rows <-> lanes access of the FFMA loops.
// pad = 4 (float64): row strides == 4 (mod 16) doubles, i.e. 32 B (mod 128 B): the DMMA A fragment (8 rows x 4
// consecutive doubles per warp load) then covers 256 B in two wavefronts.
__host__ __device__ inline SmemLayout smem_layout(int rmax, int nslots, int H, int n, int K, bool bwd,
                                                  bool fourier_coef, int ctx_reals, int pad = 1) {
  // rmax rows (points) per tile; nslots time slots per tile (shared time grid: TT <= rmax / BB, per-row grids: one
  // per row).  S1 is also the backward's per-slot G1 buffer.
  SmemLayout L;
  auto ld_of = [pad](int cols) { return pad == 1 ? cols + 1 : ((cols + 15) / 16) * 16 + 4; };
  L.ldh = ld_of(H);
  L.ldo = ld_of(2 * n);
  L.ldu = ld_of(2 * kKCU);
  size_t o = 0;
  L.H1 = o; o += (size_t)rmax * L.ldh;
  L.H2 = o; o += (size_t)rmax * L.ldh;
  L.G = o;  o += bwd ? (size_t)rmax * L.ldh : 0;
  L.S1 = o; o += (size_t)nslots * L.ldh;
  L.O = o;  o += (size_t)rmax * L.ldo;
  L.U = o;  o += (size_t)nslots * L.ldu;
  L.Coef = o; o += fourier_coef ? (size_t)nslots * n * 2 : 0;
  L.Ctx = o; o += (size_t)nslots * ctx_reals;
  L.Pm = o; o += (size_t)rmax * K;
  L.X = o;  o += (size_t)rmax * 2;
  L.total = o;
  return L;
}

// C[r][c] = sum_k A[r][k] * W[c*sc + k*sk]; rows r = lane + 32*i, columns split over warps.
template <typename R, int RP, int CB, typename Epi>
__device__ __forceinline__ void gemm_rows_simt(const R* __restrict__ sA, int lda, int Kdim, int rows,
                                               const R* __restrict__ W, int64_t sc, int64_t sk, int NC, Epi epi) {
  const int kWarps = (int)(blockDim.x >> 5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = warp * CB; c0 < NC; c0 += kWarps * CB) {
    R acc[RP][CB];
#pragma unroll
    for (int i = 0; i < RP; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) acc[i][j] = R(0);
    const R* Wc = W + (int64_t)c0 * sc;
    const int cb = min(CB, NC - c0);
    if (cb == CB) {
#pragma unroll 2
      for (int k = 0; k < Kdim; ++k) {
        R a[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) a[i] = (lane + 32 * i < rows) ? sA[(lane + 32 * i) * lda + k] : R(0);
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          R w = __ldg(Wc + j * sc + k * sk);
#pragma unroll
          for (int i = 0; i < RP; ++i) acc[i][j] = fma(a[i], w, acc[i][j]);
        }
      }
    } else {
      for (int k = 0; k < Kdim; ++k) {
        R a[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) a[i] = (lane + 32 * i < rows) ? sA[(lane + 32 * i) * lda + k] : R(0);
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          if (j < cb) {
            R w = __ldg(Wc + j * sc + k * sk);
#pragma unroll
            for (int i = 0; i < RP; ++i) acc[i][j] = fma(a[i], w, acc[i][j]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RP; ++i) {
      const int r = lane + 32 * i;
      if (r < rows) {
#pragma unroll
        for (int j = 0; j < CB; ++j)
          if (j < cb) epi(r, c0 + j, acc[i][j]);
      }
    }
  }
}

// out[j*ldo + c] += sum_{r<rows} A1[r*ld1 + j] * A2[r*ld2 + c]   (weight gradients; lanes <-> c)
template <typename R>
__device__ __forceinline__ void gemm_tn_accum_simt(const R* __restrict__ sA1, int ld1, int NJ,
                                                   const R* __restrict__ sA2, int ld2, int NC, int rows, R* out,
                                                   int64_t ldo) {
  constexpr int JB = 8, CP = 2;
  const int kWarps = (int)(blockDim.x >> 5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = warp * JB; j0 < NJ; j0 += kWarps * JB) {
    const int jb = min(JB, NJ - j0);
    for (int c0 = 0; c0 < NC; c0 += 32 * CP) {
      R acc[JB][CP];
#pragma unroll
      for (int j = 0; j < JB; ++j)
#pragma unroll
        for (int q = 0; q < CP; ++q) acc[j][q] = R(0);
      for (int r = 0; r < rows; ++r) {
        R a2[CP];
#pragma unroll
        for (int q = 0; q < CP; ++q) {
          int c = c0 + lane + 32 * q;
          a2[q] = (c < NC) ? sA2[r * ld2 + c] : R(0);
        }
#pragma unroll
        for (int j = 0; j < JB; ++j) {
          R a1 = (j < jb) ? sA1[r * ld1 + j0 + j] : R(0);
#pragma unroll
          for (int q = 0; q < CP; ++q) acc[j][q] = fma(a1, a2[q], acc[j][q]);
        }
      }
#pragma unroll
      for (int j = 0; j < JB; ++j) {
        if (j < jb) {
#pragma unroll
          for (int q = 0; q < CP; ++q) {
            int c = c0 + lane + 32 * q;
            if (c < NC) out[(int64_t)(j0 + j) * ldo + c] += acc[j][q];
          }
        }
      }
    }
  }
}


// ---- float64 on the tensor pipe ---------------------------------------------------------------------------------
// D(8x8) += A(8x4) . B(4x8), one double of A and of B and two of D per lane:
//   A[row = lane / 4][k = lane % 4],  B[k = lane % 4][col = lane / 4],  D[row = lane / 4][col = 2 (lane % 4) + {0, 1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// gemm_rows on DMMA: the tile's (at most 32) rows are four 8-row blocks, the NC columns blocks of 8; a work item is
// MTP row blocks x one column block (the B fragment -- the weights, read through L1 -- is reused MTP times) and the
// items are dealt to the warps.  Rows past `rows` hold stale shared memory: they only reach their own (discarded)
// output rows.
template <int MT, int MTP, typename Epi>
__device__ __forceinline__ void gemm_rows_dmma_t(const double* __restrict__ sA, int lda, int Kdim, int rows,
                                                 const double* __restrict__ W, int64_t sc, int64_t sk, int NC,
                                                 Epi epi) {
  const int kWarps = (int)(blockDim.x >> 5);
  constexpr int kGroups = MT / MTP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const int nb_n = (NC + 7) >> 3;
  for (int item = warp; item < nb_n * kGroups; item += kWarps) {
    const int nb = item / kGroups, mg = item - nb * kGroups;
    if (mg * MTP * 8 >= rows) continue;
    const int cb = nb * 8 + lr;
    const bool cok = cb < NC;
    const double* wp = W + (int64_t)(cok ? cb : 0) * sc;
    const int rbase = mg * MTP * 8 + lr;
    const double* ap = sA + rbase * lda;
    double acc[MTP][2];
#pragma unroll
    for (int m = 0; m < MTP; ++m) acc[m][0] = acc[m][1] = 0.0;
#pragma unroll 2
    for (int k0 = 0; k0 < Kdim; k0 += 4) {
      const int k = k0 + lk;
      const bool kok = k < Kdim;
      const double b = (cok && kok) ? __ldg(wp + (int64_t)k * sk) : 0.0;
#pragma unroll
      for (int m = 0; m < MTP; ++m) {
        // (rows past `rows` are not read: the per-slot matrices only hold `rows` of them)
        const double a = (kok && rbase + m * 8 < rows) ? ap[m * 8 * lda + k] : 0.0;
        dmma884(acc[m][0], acc[m][1], a, b);
      }
    }
    const int c = nb * 8 + 2 * lk;
#pragma unroll
    for (int m = 0; m < MTP; ++m) {
      const int r = rbase + m * 8;
      if (r < rows) {
        if (c < NC) epi(r, c, acc[m][0]);
        if (c + 1 < NC) epi(r, c + 1, acc[m][1]);
      }
    }
  }
}
// MT = 8-row blocks per tile (4 or 8)
template <int MT, typename Epi>
__device__ __forceinline__ void gemm_rows_dmma(const double* __restrict__ sA, int lda, int Kdim, int rows,
                                               const double* __restrict__ W, int64_t sc, int64_t sk, int NC, Epi epi) {
  const int kWarps = (int)(blockDim.x >> 5);
  const int nb_n = (NC + 7) >> 3;
  const int mt_used = (rows + 7) >> 3;
  // enough items for every warp, as many row blocks per item as that allows (B fragment reuse)
  if (nb_n >= kWarps || mt_used <= 1) gemm_rows_dmma_t<MT, MT>(sA, lda, Kdim, rows, W, sc, sk, NC, epi);
  else if (2 * nb_n >= kWarps || mt_used <= 2) gemm_rows_dmma_t<MT, MT / 2>(sA, lda, Kdim, rows, W, sc, sk, NC, epi);
  else gemm_rows_dmma_t<MT, MT / 4>(sA, lda, Kdim, rows, W, sc, sk, NC, epi);
}

// gemm_tn_accum on DMMA: out[j][c] += sum_r A1[r][j] A2[r][c]; work items are 8 x 8 blocks of `out`, the reduction runs
// over the tile's rows four at a time.  Every element of `out` is read-modify-written by exactly one lane (fixed
// assignment): deterministic, like the FFMA version.
__device__ __forceinline__ void gemm_tn_accum_dmma(const double* __restrict__ sA1, int ld1, int NJ,
                                                   const double* __restrict__ sA2, int ld2, int NC, int rows,
                                                   double* out, int64_t ldo) {
  const int kWarps = (int)(blockDim.x >> 5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const int jb_n = (NJ + 7) >> 3, cb_n = (NC + 7) >> 3;
  for (int item = warp; item < jb_n * cb_n; item += kWarps) {
    const int jb = item / cb_n, cb = item - jb * cb_n;
    const int j = jb * 8 + lr, c = cb * 8 + lr;
    const bool jok = j < NJ, cok = c < NC;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
    for (int r0 = 0; r0 < rows; r0 += 4) {
      const int r = r0 + lk;
      const bool rok = r < rows;
      const double a = (rok && jok) ? sA1[r * ld1 + j] : 0.0;
      const double b = (rok && cok) ? sA2[r * ld2 + c] : 0.0;
      dmma884(a0, a1, a, b);
    }
    const int co = cb * 8 + 2 * lk;
    if (jok) {
      double* o = out + (int64_t)j * ldo + co;
      if (co < NC) o[0] += a0;
      if (co + 1 < NC) o[1] += a1;
    }
  }
}

template <typename R, int RP, int CB, typename Epi>
__device__ __forceinline__ void gemm_rows(const R* __restrict__ sA, int lda, int Kdim, int rows,
                                          const R* __restrict__ W, int64_t sc, int64_t sk, int NC, Epi epi) {
  if constexpr (sizeof(R) == 8) {
    gemm_rows_dmma<4 * RP>(sA, lda, Kdim, rows, W, sc, sk, NC, epi);
  } else {
    gemm_rows_simt<R, RP, CB>(sA, lda, Kdim, rows, W, sc, sk, NC, epi);
  }
}
template <typename R>
__device__ __forceinline__ void gemm_tn_accum(const R* __restrict__ sA1, int ld1, int NJ, const R* __restrict__ sA2,
                                              int ld2, int NC, int rows, R* out, int64_t ldo) {
  if constexpr (sizeof(R) == 8) gemm_tn_accum_dmma(sA1, ld1, NJ, sA2, ld2, NC, rows, out, ldo);
  else gemm_tn_accum_simt<R>(sA1, ld1, NJ, sA2, ld2, NC, rows, out, ldo);
}

// out[c] += sum_{r<rows} A[r*ld + c]
template <typename R>
__device__ __forceinline__ void colsum_accum(const R* __restrict__ sA, int ld, int NC, int rows, R* out) {
  for (int c = threadIdx.x; c < NC; c += (int)blockDim.x) {
    R s = R(0);
    for (int r = 0; r < rows; ++r) s += sA[r * ld + c];
    out[c] += s;
  }
}

__device__ __forceinline__ void atomic_add(float* a, float v) { atomicAdd(a, v); }
__device__ __forceinline__ void atomic_add(double* a, double v) { atomicAdd(a, v); }

// per-slot context kept in shared memory
template <typename R>
struct SlotCtx {
  R t, T, gamma, inv, ratio, pref, zc, zs;
};

template <typename R, int RP, bool BWD>
__global__ void __launch_bounds__(SimtCfg<R>::kThreads, 1) fused_simt_kernel(FusedArgs<R> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* sm = reinterpret_cast<R*>(smem_raw);
  constexpr int RMAX = 32 * RP;
  // (launched with SimtCfg<R>::kThreads threads, or half of that with two CTAs per SM: SimtPlan::threads)
  const int kThreads = (int)blockDim.x, kWarps = kThreads >> 5;
  const IltParams<R> P = A.P;
  const int H = A.H, n = A.n, K = A.K, D = A.D;
  const bool dehoog = P.family == NL_DEHOOG;
  const bool fourier_coef = P.family == NL_FOURIER;
  const SmemLayout L = smem_layout(RMAX, A.t_mode == NL_T_PER_ROW ? RMAX : A.TT, H, n, K, BWD, fourier_coef,
                                   (int)(sizeof(SlotCtx<R>) / sizeof(R)), SimtCfg<R>::kPad);
  R* sH1 = sm + L.H1;
  R* sH2 = sm + L.H2;
  R* sG = sm + L.G;
  R* sS1 = sm + L.S1;
  R* sO = sm + L.O;
  R* sU = sm + L.U;
  R* sCoef = sm + L.Coef;
  SlotCtx<R>* sCtx = reinterpret_cast<SlotCtx<R>*>(sm + L.Ctx);
  R* sP = sm + L.Pm;
  R* sX = sm + L.X;   // [RMAX] accumulators / upstream grads
  R* sGp = sX + RMAX; // [RMAX] g * pref
  const int ldh = L.ldh, ldo = L.ldo, ldu = L.ldu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ldw1 = 2 * n + K;
  const bool per_row = A.t_mode == NL_T_PER_ROW;
  const int rows = A.TT * A.BB;
  const int NS = per_row ? rows : A.TT;
  const int ntiles = A.nbt * A.ntt;

  // slab sections (backward)
  R* slab = BWD ? A.slab + (int64_t)blockIdx.x * A.slab_len : nullptr;
  R* gW1 = slab;
  R* gb1 = gW1 + (int64_t)H * ldw1;
  R* gW2 = gb1 + H;
  R* gb2 = gW2 + (int64_t)H * H;
  R* gW3 = gb2 + H;
  R* gb3 = gW3 + (int64_t)2 * D * n * H;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = (tile % A.nbt) * A.BB;
    const int t0 = (tile / A.nbt) * A.TT;
    auto row_b = [&](int r) { return b0 + r % A.BB; };
    auto row_t = [&](int r) { return t0 + r / A.BB; };
    auto row_valid = [&](int r) { return r < rows && row_b(r) < A.B && row_t(r) < A.T; };
    auto row_slot = [&](int r) { return per_row ? r : r / A.BB; };

    // ---- 0. slot contexts, p rows
    for (int s = tid; s < NS; s += kThreads) {
      R tv = R(1);
      if (per_row) {
        if (row_valid(s)) tv = A.t[(int64_t)row_b(s) * A.T + row_t(s)];
      } else if (t0 + s < A.T) {
        tv = A.t[t0 + s];
      }
      TimeCtx<R> c = make_time_ctx<R>(P, tv, nullptr, 0);
      SlotCtx<R> sc;
      sc.t = c.t; sc.T = c.T; sc.gamma = c.gamma; sc.inv = c.inv; sc.ratio = c.ratio; sc.pref = c.pref;
      sc.zc = R(1); sc.zs = R(0);
      if (dehoog) m_sincos((R(kPi) * c.t) / c.T, &sc.zs, &sc.zc);
      sCtx[s] = sc;
    }
    for (int i = tid; i < rows * K; i += kThreads) {
      int r = i / K, j = i - r * K;
      sP[i] = row_valid(r) ? A.p[(int64_t)row_b(r) * K + j] : R(0);
    }
    for (int i = tid; i < NS * ldh; i += kThreads) sS1[i] = R(0);
    __syncthreads();

    // ---- 1. s points -> sphere features -> S1 = W1s . u   (chunks of kKCU terms)
    auto fill_features = [&](int kc0) {
      for (int i = tid; i < NS * kKCU; i += kThreads) {
        int s = i / kKCU, kk = i - s * kKCU, k = kc0 + kk;
        R f0 = R(0), f1 = R(0);
        if (k < n) {
          SlotCtx<R> sc = sCtx[s];
          TimeCtx<R> c{sc.t, sc.T, sc.gamma, sc.inv, sc.ratio, sc.pref};
          s_features(P, s_point(P, c, k), &f0, &f1);
          if (fourier_coef) {
            Cx<R> w = reduce_coef(P, c, k);
            sCoef[((int64_t)s * n + k) * 2] = w.re;
            sCoef[((int64_t)s * n + k) * 2 + 1] = w.im;
          }
        }
        sU[s * ldu + kk] = f0;
        sU[s * ldu + kKCU + kk] = f1;
      }
    };
    for (int kc0 = 0; kc0 < n; kc0 += kKCU) {
      fill_features(kc0);
      __syncthreads();
      const int kc = min(kKCU, n - kc0);
      gemm_rows<R, RP, 8>(sU, ldu, kc, NS, A.W1 + kc0, ldw1, 1, H,
                          [&](int r, int c, R v) { sS1[r * ldh + c] += v; });
      gemm_rows<R, RP, 8>(sU + kKCU, ldu, kc, NS, A.W1 + n + kc0, ldw1, 1, H,
                          [&](int r, int c, R v) { sS1[r * ldh + c] += v; });
      __syncthreads();
    }

    // ---- 2. h1 = tanh(S1[slot] + W1p.p + b1)
    for (int i = tid; i < rows * H; i += kThreads) {
      int r = i / H, c = i - r * H;
      R z = sS1[row_slot(r) * ldh + c] + __ldg(A.b1 + c);
      for (int j = 0; j < K; ++j) z = fma(__ldg(A.W1 + (int64_t)c * ldw1 + 2 * n + j), sP[r * K + j], z);
      sH1[r * ldh + c] = m_tanh(z);
    }
    __syncthreads();
    // ---- 3. h2 = tanh(W2.h1 + b2)
    gemm_rows<R, RP, 8>(sH1, ldh, H, rows, A.W2, H, 1, H,
                        [&](int r, int c, R v) { sH2[r * ldh + c] = m_tanh(v + __ldg(A.b2 + c)); });
    if (BWD) {
      for (int i = tid; i < rows * ldh; i += kThreads) sG[i] = R(0);
    }
    __syncthreads();

    // ---- 4. per d_obs: layer 3, heads, reduction over terms
    for (int d = 0; d < D; ++d) {
      const int ja = d * n, jb = (D + d) * n;
      gemm_rows<R, RP, 8>(sH2, ldh, H, rows, A.W3 + (int64_t)ja * H, H, 1, n,
                          [&](int r, int c, R v) { sO[r * ldo + c] = v + __ldg(A.b3 + ja + c); });
      gemm_rows<R, RP, 8>(sH2, ldh, H, rows, A.W3 + (int64_t)jb * H, H, 1, n,
                          [&](int r, int c, R v) { sO[r * ldo + n + c] = v + __ldg(A.b3 + jb + c); });
      for (int r = tid; r < RMAX; r += kThreads) {
        sX[r] = R(0);
        if (BWD) {
          R g = row_valid(r) ? A.gx[((int64_t)row_b(r) * A.T + row_t(r)) * D + d] : R(0);
          sGp[r] = g * sCtx[row_slot(r < rows ? r : 0)].pref;
        }
      }
      __syncthreads();

      if (!dehoog) {
        R acc[RP];
#pragma unroll
        for (int i = 0; i < RP; ++i) acc[i] = R(0);
        for (int k = warp; k < n; k += kWarps) {
#pragma unroll
          for (int i = 0; i < RP; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) {
              HeadOut<R> h = head_forward<R>(P.head, sO[r * ldo + k], sO[r * ldo + n + k]);
              Cx<R> w;
              if (fourier_coef) {
                const int s = row_slot(r);
                w.re = sCoef[((int64_t)s * n + k) * 2];
                w.im = sCoef[((int64_t)s * n + k) * 2 + 1];
              } else {
                w.re = __ldg(P.eta + 2 * k);
                w.im = __ldg(P.eta + 2 * k + 1);
              }
              if (!BWD) {
                acc[i] += h.fre * w.re - h.fim * w.im;
              } else {
                R g = sGp[r], goa, gob;
                head_backward<R>(P.head, h, g * w.re, -g * w.im, &goa, &gob);
                sO[r * ldo + k] = goa;
                sO[r * ldo + n + k] = gob;
              }
            }
          }
        }
        if (!BWD) {
#pragma unroll
          for (int i = 0; i < RP; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) atomic_add(&sX[r], acc[i]);
          }
        }
      } else {
        // De Hoog: the quotient-difference recurrence is not run here (one CTA per SM and <= 128 rows per tile
        // would leave it on a few warps): forward = F out to its [d][k][t][b] plane, backward = dL/dF in.
        const int64_t plane = (int64_t)A.T * A.B;
        for (int k = warp; k < n; k += kWarps) {
#pragma unroll
          for (int i = 0; i < RP; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) {
              HeadOut<R> h = head_forward<R>(P.head, sO[r * ldo + k], sO[r * ldo + n + k]);
              const int64_t idx = ((int64_t)d * n + k) * plane + (int64_t)row_t(r) * A.B + row_b(r);
              if (!BWD) {
                if (row_valid(r)) A.fplane[idx] = Cx<R>{h.fre, h.fim};
              } else {
                Cx<R> df{R(0), R(0)};
                if (row_valid(r)) df = A.dfplane[idx];
                R goa, gob;
                head_backward<R>(P.head, h, df.re, df.im, &goa, &gob);
                sO[r * ldo + k] = goa;
                sO[r * ldo + n + k] = gob;
              }
            }
          }
        }
      }
      __syncthreads();

      if (!BWD) {
        if (!dehoog) {  // (De Hoog: x is written by dehoog_qd_kernel)
          for (int r = tid; r < rows; r += kThreads) {
            if (row_valid(r)) A.x[((int64_t)row_b(r) * A.T + row_t(r)) * D + d] = sCtx[row_slot(r)].pref * sX[r];
          }
        }
      } else {
        // dH2 += dO . W3[rows of d]   ;   dW3 / db3 slabs
        gemm_rows<R, RP, 8>(sO, ldo, n, rows, A.W3 + (int64_t)ja * H, 1, H, H,
                            [&](int r, int c, R v) { sG[r * ldh + c] += v; });
        gemm_rows<R, RP, 8>(sO + n, ldo, n, rows, A.W3 + (int64_t)jb * H, 1, H, H,
                            [&](int r, int c, R v) { sG[r * ldh + c] += v; });
        gemm_tn_accum<R>(sO, ldo, n, sH2, ldh, H, rows, gW3 + (int64_t)ja * H, H);
        gemm_tn_accum<R>(sO + n, ldo, n, sH2, ldh, H, rows, gW3 + (int64_t)jb * H, H);
        colsum_accum<R>(sO, ldo, n, rows, gb3 + ja);
        colsum_accum<R>(sO + n, ldo, n, rows, gb3 + jb);
      }
      __syncthreads();
    }  // d

    if (BWD) {
      // ---- 5. layer 2 / layer 1 backward
      for (int i = tid; i < rows * H; i += kThreads) {
        int r = i / H, c = i - r * H;
        R h2 = sH2[r * ldh + c];
        sG[r * ldh + c] *= (R(1) - h2 * h2);
      }
      __syncthreads();
      gemm_tn_accum<R>(sG, ldh, H, sH1, ldh, H, rows, gW2, H);
      colsum_accum<R>(sG, ldh, H, rows, gb2);
      // dz1 -> sH2 (h2 is dead)
      gemm_rows<R, RP, 8>(sG, ldh, H, rows, A.W2, 1, H, H, [&](int r, int c, R v) {
        R h1 = sH1[r * ldh + c];
        sH2[r * ldh + c] = v * (R(1) - h1 * h1);
      });
      __syncthreads();
      R* sZ1 = sH2;
      colsum_accum<R>(sZ1, ldh, H, rows, gb1);
      gemm_tn_accum<R>(sZ1, ldh, H, sP, K, K, rows, gW1 + 2 * n, ldw1);
      for (int i = tid; i < rows * K; i += kThreads) {
        int r = i / K, j = i - r * K;
        if (row_valid(r)) {
          R s = R(0);
          for (int c = 0; c < H; ++c) s = fma(sZ1[r * ldh + c], __ldg(A.W1 + (int64_t)c * ldw1 + 2 * n + j), s);
          atomic_add(A.dp + (int64_t)row_b(r) * K + j, s);
        }
      }
      // G1[slot] = sum over the rows of the slot
      const R* sG1 = sZ1;
      if (!per_row) {
        for (int i = tid; i < NS * H; i += kThreads) {
          int s = i / H, c = i - s * H;
          R acc = R(0);
          for (int bb = 0; bb < A.BB; ++bb) acc += sZ1[(s * A.BB + bb) * ldh + c];
          sS1[s * ldh + c] = acc;
        }
        sG1 = sS1;
      }
      __syncthreads();
      for (int kc0 = 0; kc0 < n; kc0 += kKCU) {
        fill_features(kc0);
        __syncthreads();
        const int kc = min(kKCU, n - kc0);
        gemm_tn_accum<R>(sG1, ldh, H, sU, ldu, kc, NS, gW1 + kc0, ldw1);
        gemm_tn_accum<R>(sG1, ldh, H, sU + kKCU, ldu, kc, NS, gW1 + n + kc0, ldw1);
        __syncthreads();
      }
    }
    __syncthreads();
  }  // tiles
}

// grads[i] = sum over CTA slabs (deterministic order)
template <typename R>
__global__ void reduce_slabs_kernel(const R* __restrict__ slab, int64_t slab_len, int nslabs, R* gW1, R* gb1, R* gW2,
                                    R* gb2, R* gW3, R* gb3, int64_t nW1, int64_t nb1, int64_t nW2, int64_t nb2,
                                    int64_t nW3, int64_t nb3) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= slab_len) return;
  R s = R(0);
  for (int c = 0; c < nslabs; ++c) s += slab[(int64_t)c * slab_len + i];
  int64_t o = i;
  if (o < nW1) { gW1[o] = s; return; }
  o -= nW1;
  if (o < nb1) { gb1[o] = s; return; }
  o -= nb1;
  if (o < nW2) { gW2[o] = s; return; }
  o -= nW2;
  if (o < nb2) { gb2[o] = s; return; }
  o -= nb2;
  if (o < nW3) { gW3[o] = s; return; }
  o -= nW3;
  if (o < nb3) gb3[o] = s;
}

// ---------------------------------------------------------------------------------------
int device_sm_count() {
  // per device: a process may drive several GPUs (the caller makes the tensors' device current)
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

SimtPlan simt_plan(const nl_desc* d, int pass) {
  SimtPlan pl{};
  const bool bwd = pass != 0;
  const size_t rs = d->dtype == NL_F64 ? 8 : 4;
  if (d->H < 1 || d->H > 256 || d->K < 1 || d->K > 64 || d->n < 1 || d->D < 1) return pl;
  if (d->family == NL_DEHOOG && (d->n < 3 || d->n % 2 == 0)) return pl;
  const int ctx_reals = 8;
  const size_t limit = 227 * 1024;
  int rp = 0;
  // rows per tile: 64 when the shared-memory matrices fit (float64: shared time grids only -- the per-slot matrices
  // then hold rmax / BB slots instead of rmax), else 32
  for (int cand = 2; cand >= 1; --cand) {
    const int rmax_c = 32 * cand;
    const int BBc = std::max(1, std::min(d->B, rmax_c));
    const int TTc = std::max(1, std::min(d->T, rmax_c / BBc));
    const int nslots = d->t_mode == NL_T_PER_ROW ? rmax_c : TTc;
    SmemLayout L = smem_layout(rmax_c, nslots, d->H, d->n, d->K, bwd, d->family == NL_FOURIER, ctx_reals,
                               d->dtype == NL_F64 ? SimtCfg<double>::kPad : SimtCfg<float>::kPad);
    if (L.total * rs <= limit) {
      rp = cand;
      pl.smem = L.total * rs;
      break;
    }
  }
  if (rp == 0) return pl;
  pl.threads = d->dtype == NL_F64 ? SimtCfg<double>::kThreads : SimtCfg<float>::kThreads;
  int ctas = 1;
  if (d->dtype == NL_F64) {
    // float64 FORWARD: two CTAs of 256 threads and 32-row tiles per SM instead of one of 512 threads and 64 rows when
    // the shared memory allows it (same warps and rows in flight, but the phases of the two CTAs -- GEMMs on the DMMA
    // pipe, float64 transcendentals -- overlap instead of alternating): 3.56 -> 3.35 ms at B=1024 x T=1024.  The
    // backward is slower that way (10.1 -> 12.1 ms: twice the slab read-modify-write passes per point) and keeps one.
    // NL_B200_F64_CTAS: 1 = never, 2 = forward only (default), 3 = both passes.
    static const int want2 = [] { const char* v = getenv("NL_B200_F64_CTAS"); return v ? atoi(v) : 2; }();
    const int BB1 = std::max(1, std::min(d->B, 32)), TT1 = std::max(1, std::min(d->T, 32 / BB1));
    SmemLayout L1 = smem_layout(32, d->t_mode == NL_T_PER_ROW ? 32 : TT1, d->H, d->n, d->K, bwd,
                                d->family == NL_FOURIER, ctx_reals, SimtCfg<double>::kPad);
    if ((want2 == 3 || (want2 == 2 && !bwd)) && (L1.total * rs + 1024) * 2 <= (size_t)228 * 1024) {
      rp = 1;
      pl.smem = L1.total * rs;
      pl.threads = SimtCfg<double>::kThreads / 2;
      ctas = 2;
    }
  }
  const int rmax = 32 * rp;
  pl.rp = rp;
  pl.BB = std::max(1, std::min(d->B, rmax));
  pl.TT = std::max(1, std::min(d->T, rmax / pl.BB));
  const int64_t nbt = (d->B + pl.BB - 1) / pl.BB, ntt = (d->T + pl.TT - 1) / pl.TT;
  const int64_t ntiles = nbt * ntt;
  pl.grid = (int)std::min<int64_t>(ntiles, (int64_t)device_sm_count() * ctas);
  if (pl.grid < 1) pl.grid = 1;
  const int64_t ldw1 = 2 * d->n + d->K;
  pl.slab_len = (size_t)d->H * ldw1 + d->H + (size_t)d->H * d->H + d->H + (size_t)2 * d->D * d->n * d->H +
                (size_t)2 * d->D * d->n;
  size_t ws = 0;
  if (bwd) ws += align_up(pl.slab_len * rs * pl.grid, 256);
  if (d->family == NL_DEHOOG)  // F plane (+ dL/dF plane) + what the recurrence kernels need
    ws += (bwd ? 2 : 1) * align_up((size_t)d->D * d->n * (size_t)d->T * d->B * 2 * rs, 256) +
          align_up(dehoog_qd_scratch_bytes(d, bwd), 256);
  pl.ws_bytes = ws + 256;
  pl.ok = 1;
  return pl;
}

template <typename R, int RP>
static cudaError_t launch_rp(const FusedArgs<R>& A, bool bwd, int grid, int threads, size_t smem, cudaStream_t st) {
  cudaError_t e;
  if (bwd) {
    e = cudaFuncSetAttribute(fused_simt_kernel<R, RP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fused_simt_kernel<R, RP, true><<<grid, threads, smem, st>>>(A);
  } else {
    e = cudaFuncSetAttribute(fused_simt_kernel<R, RP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fused_simt_kernel<R, RP, false><<<grid, threads, smem, st>>>(A);
  }
  return cudaGetLastError();
}

template <typename R>
cudaError_t launch_fused_simt(const nl_desc* d, bool bwd, const void* p, const void* t, const void* const* weights,
                              const void* beta, const void* eta, void* x, const void* gx, void* const* grads,
                              void* ws, cudaStream_t st, int* launches) {
  SimtPlan pl = simt_plan(d, bwd ? 1 : 0);
  if (!pl.ok) return cudaErrorInvalidValue;
  if ((int64_t)d->B * d->T == 0) {
    if (bwd) {  // empty batch: every gradient is zero
      const int64_t ldw1e = 2 * d->n + d->K;
      const size_t cnt[7] = {(size_t)d->H * ldw1e, (size_t)d->H, (size_t)d->H * d->H, (size_t)d->H,
                             (size_t)2 * d->D * d->n * d->H, (size_t)2 * d->D * d->n, (size_t)d->B * d->K};
      for (int i = 0; i < 7; ++i) {
        if (cnt[i] == 0) continue;
        cudaError_t e0 = cudaMemsetAsync(grads[i], 0, cnt[i] * sizeof(R), st);
        if (e0 != cudaSuccess) return e0;
      }
    }
    return cudaSuccess;
  }
  FusedArgs<R> A{};
  A.P.family = d->family;
  A.P.n = d->n;
  A.P.head = d->head;
  A.P.alpha = (R)d->alpha;
  A.P.log_tol = (R)d->log_tol;
  A.P.scale = (R)d->scale;
  A.P.beta = (const R*)beta;
  A.P.eta = (const R*)eta;
  A.B = d->B; A.T = d->T; A.K = d->K; A.H = d->H; A.D = d->D; A.n = d->n; A.t_mode = d->t_mode;
  A.BB = pl.BB; A.TT = pl.TT;
  A.nbt = (d->B + pl.BB - 1) / pl.BB;
  A.ntt = (d->T + pl.TT - 1) / pl.TT;
  A.p = (const R*)p; A.t = (const R*)t;
  A.W1 = (const R*)weights[0]; A.b1 = (const R*)weights[1]; A.W2 = (const R*)weights[2];
  A.b2 = (const R*)weights[3]; A.W3 = (const R*)weights[4]; A.b3 = (const R*)weights[5];
  A.x = (R*)x; A.gx = (const R*)gx;
  unsigned char* w = (unsigned char*)ws;
  w = (unsigned char*)align_up((size_t)w, 256);
  cudaError_t e;
  if (bwd) {
    A.slab = (R*)w;
    A.slab_len = (int64_t)pl.slab_len;
    size_t bytes = align_up(pl.slab_len * sizeof(R) * pl.grid, 256);
    e = cudaMemsetAsync(w, 0, bytes, st);
    if (e != cudaSuccess) return e;
    w += bytes;
    A.dp = (R*)grads[6];
    e = cudaMemsetAsync(A.dp, 0, (size_t)d->B * d->K * sizeof(R), st);
    if (e != cudaSuccess) return e;
  }
  auto run = [&](const FusedArgs<R>& args, bool backward) {
    return pl.rp == 2 ? launch_rp<R, 2>(args, backward, pl.grid, pl.threads, pl.smem, st)
                      : launch_rp<R, 1>(args, backward, pl.grid, pl.threads, pl.smem, st);
  };
  if (d->family == NL_DEHOOG) {
    // forward: MLP -> F plane, recurrence -> x.   backward: MLP -> F plane (recomputed), recurrence forward +
    // reverse sweep -> dL/dF plane, then the backward kernel proper with dL/dF in place of grad_x * coefficients.
    const size_t fbytes = align_up((size_t)d->D * d->n * (size_t)d->T * d->B * 2 * sizeof(R), 256);
    Cx<R>* F = (Cx<R>*)w;
    w += fbytes;
    Cx<R>* dF = nullptr;
    if (bwd) {
      dF = (Cx<R>*)w;
      w += fbytes;
    }
    FusedArgs<R> Af = A;
    Af.fplane = F;
    if (bwd) {  // the forward instantiation has no slab / dp
      Af.slab = nullptr;
      Af.dp = nullptr;
      const SimtPlan plf = simt_plan(d, 0);  // (the forward plan may pick a different tile geometry)
      if (!plf.ok) return cudaErrorInvalidValue;
      Af.BB = plf.BB; Af.TT = plf.TT;
      Af.nbt = (d->B + plf.BB - 1) / plf.BB;
      Af.ntt = (d->T + plf.TT - 1) / plf.TT;
      e = plf.rp == 2 ? launch_rp<R, 2>(Af, false, plf.grid, plf.threads, plf.smem, st)
                      : launch_rp<R, 1>(Af, false, plf.grid, plf.threads, plf.smem, st);
    } else {
      e = run(Af, false);
    }
    if (e != cudaSuccess) return e;
    e = launch_dehoog_qd(d, bwd, t, F, gx, x, dF, w, st);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 2;
    if (!bwd) return cudaSuccess;
    A.dfplane = dF;
  }
  e = run(A, bwd);
  if (e != cudaSuccess) return e;
  if (launches) *launches += 1;
  if (bwd) {
    const int64_t ldw1 = 2 * d->n + d->K;
    const int64_t nW1 = (int64_t)d->H * ldw1, nb1 = d->H, nW2 = (int64_t)d->H * d->H, nb2 = d->H,
                  nW3 = (int64_t)2 * d->D * d->n * d->H, nb3 = (int64_t)2 * d->D * d->n;
    int threads = 256;
    int64_t blocks = ((int64_t)pl.slab_len + threads - 1) / threads;
    reduce_slabs_kernel<R><<<(unsigned)blocks, threads, 0, st>>>(A.slab, A.slab_len, pl.grid, (R*)grads[0],
                                                               (R*)grads[1], (R*)grads[2], (R*)grads[3],
                                                               (R*)grads[4], (R*)grads[5], nW1, nb1, nW2, nb2, nW3,
                                                               nb3);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
  }
  return cudaSuccess;
}

cudaError_t launch_reduce_slabs_f32(const float* slab, long long slab_len, int nslabs, void* const* grads,
                                    const nl_desc* d, cudaStream_t st) {
  const int64_t ldw1 = 2 * d->n + d->K;
  const int64_t nW1 = (int64_t)d->H * ldw1, nb1 = d->H, nW2 = (int64_t)d->H * d->H, nb2 = d->H,
                nW3 = (int64_t)2 * d->D * d->n * d->H, nb3 = (int64_t)2 * d->D * d->n;
  int threads = 256;
  int64_t blocks = (slab_len + threads - 1) / threads;
  reduce_slabs_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(slab, slab_len, nslabs, (float*)grads[0],
                                                                 (float*)grads[1], (float*)grads[2],
                                                                 (float*)grads[3], (float*)grads[4],
                                                                 (float*)grads[5], nW1, nb1, nW2, nb2, nW3, nb3);
  return cudaGetLastError();
}

template cudaError_t launch_fused_simt<float>(const nl_desc*, bool, const void*, const void*, const void* const*,
                                              const void*, const void*, void*, const void*, void* const*, void*,
                                              cudaStream_t, int*);
template cudaError_t launch_fused_simt<double>(const nl_desc*, bool, const void*, const void*, const void* const*,
                                               const void*, const void*, void*, const void*, void* const*, void*,
                                               cudaStream_t, int*);

}  // namespace nl
