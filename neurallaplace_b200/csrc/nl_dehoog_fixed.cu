// Fixed-contour De Hoog (reference: DeHoog.fixed_line_integrate,
// torchlaplace/inverse_laplace.py:587-701).  ONE set of s points (contour scale T = scale * time_max) serves
// every time point, so the O(M^2) quotient-difference table is built ONCE per contour row and only the O(M)
// Pade recurrence with z = exp(i pi t / T) runs per time point -- against the general line integral, which
// rebuilds the table for every (row, t) (what `forward(fs, t, time_max=...)` costs in the reference, l.897-1042).
//
//   forward : qd_rows_kernel<false> (thread per contour row: a_k -> d_0..d_2M)
//             pade_kernel<false>    (thread per (row, t): x = exp(gamma t)/T * Re(A*/B*))
//   backward: qd_rows_kernel<true>  (table kept)
//             pade_kernel<true>     (thread per (row, t): dw/dd_i; g-weighted sums over each tile of 128 time
//                                    points by warp shuffles + shared memory -> one partial per tile, no atomics)
//             qd_reverse_kernel     (thread per row: partials summed in tile order, reverse sweep -> dL/dF)
// The arithmetic is the per-element recurrence of nl_dehoog.cuh, split at d (dehoog_qd_* / dehoog_pade*).
#include <algorithm>

#include "nl_dehoog.cuh"
#include "nl_kernels.h"

namespace nl {
namespace {

constexpr int kTile = 128;  // time points per tile = threads per CTA of the Pade kernels

template <typename R>
struct FixedArgs {
  int64_t rows, T;
  int n, M, LD;
  int64_t tiles_per_row;  // ceil(T / kTile)
  R contour_T, gamma;
  const Cx<R>* F;     // [rows][n]
  const R* t;         // [T]
  const R* gx;        // [rows][T]
  R* x;               // [rows][T]
  Cx<R>* dF;          // [rows][n]
  Cx<R>* tab;         // per row: forward 3 LD (d at 2 LD), backward (2M+2) LD (d at (2M+1) LD); stride = rows
  Cx<R>* thr;         // per launched Pade thread: dbar (LD) + A/B history (2 LD); stride = thr_stride
  int64_t thr_stride;
  Cx<R>* part;        // [rows][tiles_per_row][LD] partial sums of g * dw/dd_i
  Cx<R>* adj;         // per row: summed dbar (LD) + ping-pong adjoint columns (4 LD); stride = rows
};

template <typename R, bool KEEP>
__global__ void qd_rows_kernel(FixedArgs<R> A) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.rows) return;
  const Cx<R>* Fp = A.F + row * A.n;
  auto a = [&](int k) { return Fp[k]; };
  Scratch<R> tab{A.tab, A.rows, row};
  if (KEEP) dehoog_qd_store<R>(A.M, a, tab);
  else dehoog_qd_coeffs<R>(A.M, a, tab);
}

template <typename R, bool BWD>
__global__ void __launch_bounds__(kTile) pade_kernel(FixedArgs<R> A) {
  __shared__ Cx<R> warp_sum[kTile / 32];
  const int64_t tiles = A.rows * A.tiles_per_row;
  const int d_off = BWD ? (2 * A.M + 1) * A.LD : 2 * A.LD;
  const int64_t slot = (int64_t)blockIdx.x * kTile + threadIdx.x;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row = tile / A.tiles_per_row;
    const int64_t ti = (tile - row * A.tiles_per_row) * kTile + threadIdx.x;
    const bool live = ti < A.T;
    const Scratch<R> dd = Scratch<R>{A.tab, A.rows, row}.at(d_off);
    auto d = [&](int i) { return dd[i]; };
    R g = R(0);
    Scratch<R> dbar{A.thr, A.thr_stride, slot};
    if (live) {
      const R t = A.t[ti];
      R sn, cs;
      m_sincos((R(kPi) * t) / A.contour_T, &sn, &cs);  // z = exp(i pi t / T), inverse_laplace.py:677
      const Cx<R> z{cs, sn};
      const R pref = m_exp(A.gamma * t) / A.contour_T;  // l.701
      if (!BWD) {
        A.x[row * A.T + ti] = pref * dehoog_pade<R>(A.M, z, d).re;
      } else {
        g = A.gx[row * A.T + ti] * pref;
        dehoog_pade_fwd_bwd<R>(A.M, z, d, dbar, dbar.at(A.LD));
      }
    }
    if (BWD) {
      // x = pref Re(w): sum_t g_t dw_t/dd_i per tile, fixed order (lanes by shuffle tree, warps 0..3 in turn)
      Cx<R>* out = A.part + tile * A.LD;
      for (int i = 0; i <= 2 * A.M; ++i) {
        Cx<R> v = live ? cscale(dbar[i], g) : Cx<R>{R(0), R(0)};
        for (int o = 16; o > 0; o >>= 1) {
          v.re += __shfl_down_sync(0xffffffffu, v.re, o);
          v.im += __shfl_down_sync(0xffffffffu, v.im, o);
        }
        if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
          Cx<R> s = warp_sum[0];
          for (int w = 1; w < kTile / 32; ++w) s = cadd(s, warp_sum[w]);
          out[i] = s;
        }
        __syncthreads();
      }
    }
  }
}

template <typename R>
__global__ void qd_reverse_kernel(FixedArgs<R> A) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.rows) return;
  const Cx<R>* Fp = A.F + row * A.n;
  Cx<R>* dFp = A.dF + row * A.n;
  auto a = [&](int k) { return Fp[k]; };
  Scratch<R> adj{A.adj, A.rows, row};
  for (int i = 0; i <= 2 * A.M; ++i) {
    Cx<R> s{R(0), R(0)};
    const Cx<R>* pp = A.part + row * A.tiles_per_row * A.LD + i;
    for (int64_t c = 0; c < A.tiles_per_row; ++c) s = cadd(s, pp[c * A.LD]);
    adj[i] = s;
  }
  // holomorphic adjoints G_k = sum_i Dbar_i dd_i/da_k;  dL/dRe a_k = Re G_k, dL/dIm a_k = -Im G_k
  auto emit = [&](int k, Cx<R> G) { dFp[k] = Cx<R>{G.re, -G.im}; };
  dehoog_qd_reverse<R>(A.M, a, emit, Scratch<R>{A.tab, A.rows, row}, adj, adj.at(A.LD));
}

struct FixedPlan {
  int M, LD;
  int64_t tiles_per_row, pade_blocks;
  size_t tab, thr, part, adj;  // complex entries
};
FixedPlan fixed_plan(int64_t rows, int n, int64_t T, bool bwd) {
  FixedPlan p{};
  p.M = (n - 1) / 2;
  p.LD = 2 * p.M + 2;
  p.tiles_per_row = (T + kTile - 1) / kTile;
  p.pade_blocks = std::min<int64_t>(rows * p.tiles_per_row, (int64_t)device_sm_count() * 16);
  p.tab = (size_t)rows * (bwd ? (size_t)(2 * p.M + 2) * p.LD : (size_t)3 * p.LD);
  p.thr = bwd ? (size_t)p.pade_blocks * kTile * 3 * p.LD : 0;
  p.part = bwd ? (size_t)rows * p.tiles_per_row * p.LD : 0;
  p.adj = bwd ? (size_t)rows * 5 * p.LD : 0;
  return p;
}

template <typename R>
cudaError_t run(int64_t rows, int n, int64_t T, bool bwd, const void* F, const void* t, double contour_T,
                double gamma, const void* gx, void* x, void* dF, void* scratch, cudaStream_t st) {
  const FixedPlan p = fixed_plan(rows, n, T, bwd);
  FixedArgs<R> A{};
  A.rows = rows; A.T = T; A.n = n; A.M = p.M; A.LD = p.LD;
  A.tiles_per_row = p.tiles_per_row;
  A.contour_T = (R)contour_T;
  A.gamma = (R)gamma;
  A.F = (const Cx<R>*)F; A.t = (const R*)t; A.gx = (const R*)gx; A.x = (R*)x; A.dF = (Cx<R>*)dF;
  Cx<R>* base = (Cx<R>*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  A.tab = base;
  A.thr = A.tab + p.tab;
  A.thr_stride = p.pade_blocks * kTile;
  A.part = A.thr + p.thr;
  A.adj = A.part + p.part;
  const unsigned row_blocks = (unsigned)((rows + 63) / 64);
  if (!bwd) {
    qd_rows_kernel<R, false><<<row_blocks, 64, 0, st>>>(A);
    pade_kernel<R, false><<<(unsigned)p.pade_blocks, kTile, 0, st>>>(A);
  } else {
    qd_rows_kernel<R, true><<<row_blocks, 64, 0, st>>>(A);
    pade_kernel<R, true><<<(unsigned)p.pade_blocks, kTile, 0, st>>>(A);
    qd_reverse_kernel<R><<<row_blocks, 64, 0, st>>>(A);
  }
  return cudaGetLastError();
}
}  // namespace

size_t dehoog_fixed_scratch_bytes(int dtype, int64_t rows, int n, int64_t T, bool bwd) {
  const FixedPlan p = fixed_plan(rows, n, T, bwd);
  return (p.tab + p.thr + p.part + p.adj) * (dtype == NL_F64 ? 16 : 8) + 512;
}

cudaError_t launch_dehoog_fixed(int dtype, int64_t rows, int n, int64_t T, bool bwd, const void* F, const void* t,
                                double contour_T, double gamma, const void* gx, void* x, void* dF, void* scratch,
                                cudaStream_t st) {
  return dtype == NL_F64 ? run<double>(rows, n, T, bwd, F, t, contour_T, gamma, gx, x, dF, scratch, st)
                         : run<float>(rows, n, T, bwd, F, t, contour_T, gamma, gx, x, dF, scratch, st);
}

}  // namespace nl
