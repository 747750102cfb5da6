// De Hoog quotient-difference recurrence as its own kernels, between the kernels that evaluate the MLP
// (reference: torchlaplace/inverse_laplace.py:737-866; the recurrence itself is nl_dehoog_sm.cuh / nl_dehoog.cuh).
// One thread per (b, t, d) element, as BASELINE.json's north_star asks ("a per-thread recurrence that stays
// parallel over batch and time"): the fused kernels run one CTA per SM with at most 128 rows per tile, which left
// the recurrence on 2-4 warps per SM; here it runs at full occupancy.  F and dL/dF are [d][k][t][b] complex
// planes, so the 32 threads of a warp (consecutive b) read and write contiguous bytes per term.
//   forward : F -> x = pref * Re(A*/B*)
//   backward: F, grad_x -> dL/dF as (g Re dw, -g Im dw), g = grad_x * pref: what head_backward takes
#include <algorithm>
#include <cstdlib>

#include "nl_dehoog.cuh"
#include "nl_dehoog_sm.cuh"
#include "nl_kernels.h"

namespace nl {

namespace {

template <typename R>
struct QdArgs {
  IltParams<R> P;
  int B, T, D, per_row;
  const R* t;
  const R* Tper;     // optional per-time-point contour scale (line_integrate_all_multi), indexed like t
  const Cx<R>* F;
  const R* gx;
  R* x;
  Cx<R>* dF;
  Cx<R>* scratch;    // per-thread table (strip variant) or the whole working set (generic variant)
  int64_t stride;    // launched threads
};

template <typename R>
struct QdItem {
  int d, ti, b;
  int64_t tb, xi;
  TimeCtx<R> c;
  Cx<R> z;
};
template <typename R>
__device__ __forceinline__ QdItem<R> qd_item(const QdArgs<R>& A, int64_t item) {
  QdItem<R> q;
  const int64_t plane = (int64_t)A.T * A.B;
  q.d = (int)(item / plane);
  q.tb = item - (int64_t)q.d * plane;
  q.ti = (int)(q.tb / A.B);
  q.b = (int)(q.tb - (int64_t)q.ti * A.B);
  const int64_t tidx = A.per_row ? (int64_t)q.b * A.T + q.ti : q.ti;
  q.c = make_time_ctx<R>(A.P, A.t[tidx], A.Tper, tidx);
  R sn, cs;
  m_sincos((R(kPi) * q.c.t) / q.c.T, &sn, &cs);  // z = exp(i*pi*t/T), inverse_laplace.py:842
  q.z = Cx<R>{cs, sn};
  q.xi = ((int64_t)q.b * A.T + q.ti) * A.D + q.d;
  return q;
}

// ---- generic variant: any number of terms, working set in global memory (nl_dehoog.cuh) ----
constexpr int kQdThreads = 128;
template <typename R, bool BWD>
__global__ void __launch_bounds__(kQdThreads) dehoog_qd_kernel(QdArgs<R> A) {
  const int n = A.P.n, M = (n - 1) / 2;
  const int64_t plane = (int64_t)A.T * A.B;
  const int64_t items = plane * A.D;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t item = slot; item < items; item += (int64_t)gridDim.x * blockDim.x) {
    const QdItem<R> q = qd_item(A, item);
    const Cx<R>* Fp = A.F + (int64_t)q.d * n * plane + q.tb;
    auto a = [&](int k) { return Fp[(int64_t)k * plane]; };
    Scratch<R> scr{A.scratch, A.stride, slot};
    if (!BWD) {
      const Cx<R> w = dehoog_forward<R>(M, q.z, a, scr);
      A.x[q.xi] = q.c.pref * w.re;
    } else {
      const R g = A.gx[q.xi] * q.c.pref;
      Cx<R>* dFp = A.dF + (int64_t)q.d * n * plane + q.tb;
      auto emit = [&](int k, Cx<R> dw) { dFp[(int64_t)k * plane] = Cx<R>{g * dw.re, -g * dw.im}; };
      dehoog_forward_backward<R>(M, q.z, a, emit, scr);
    }
  }
}

// ---- shared-memory-strip variant (nl_dehoog_sm.cuh): working columns in shared memory, table in global ----
template <typename R, bool BWD, int THREADS>
__global__ void __launch_bounds__(THREADS) dehoog_qd_sm_kernel(QdArgs<R> A) {
  extern __shared__ __align__(16) unsigned char qd_smem[];
  Cx<R>* strip = reinterpret_cast<Cx<R>*>(qd_smem) + threadIdx.x;
  const int n = A.P.n, M = (n - 1) / 2;
  const int64_t plane = (int64_t)A.T * A.B;
  const int64_t items = plane * A.D;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  char* tabp = reinterpret_cast<char*>(A.scratch + slot);
  // byte stride between table entries of one thread: < 2^32 (at most a few hundred thousand resident threads), so
  // an entry address is ONE 32x32->64 multiply-add
  const uint32_t sb = (uint32_t)(A.stride * (int64_t)sizeof(Cx<R>));
  auto col = [&](int i) -> Cx<R>& { return strip[i * THREADS]; };
  auto tab = [&](int i) -> Cx<R>& { return *reinterpret_cast<Cx<R>*>(tabp + (uint64_t)(uint32_t)i * sb); };
  auto pf = [&](int i) { asm volatile("prefetch.global.L2 [%0];" ::"l"(tabp + (uint64_t)(uint32_t)i * sb)); };
  for (int64_t item = slot; item < items; item += (int64_t)gridDim.x * blockDim.x) {
    const QdItem<R> q = qd_item(A, item);
    const Cx<R>* Fp = A.F + (int64_t)q.d * n * plane + q.tb;
    auto a = [&](int k) { return Fp[(int64_t)k * plane]; };
    auto pfa = [&](int k) { asm volatile("prefetch.global.L2 [%0];" ::"l"(Fp + (int64_t)k * plane)); };
    if (!BWD) {
      auto emit = [&](int, Cx<R>) {};
      const Cx<R> w = dehoog_sm<R, false>(M, q.z, a, emit, col, tab, pf, pfa);
      A.x[q.xi] = q.c.pref * w.re;
    } else {
      const R g = A.gx[q.xi] * q.c.pref;
      Cx<R>* dFp = A.dF + (int64_t)q.d * n * plane + q.tb;
      auto emit = [&](int k, Cx<R> dw) { dFp[(int64_t)k * plane] = Cx<R>{g * dw.re, -g * dw.im}; };
      dehoog_sm<R, true>(M, q.z, a, emit, col, tab, pf, pfa);
    }
  }
}

// threads per CTA of the strip variant: the largest of 128 / 64 / 32 whose strips (2n complex per thread) let
// three CTAs share an SM; 0 = use the generic variant
int sm_threads(const nl_desc* d) {
  static const bool off = [] { const char* v = getenv("NL_B200_DEHOOG_SM"); return v && atoi(v) == 0; }();
  if (off) return 0;
  const size_t per_thread = (size_t)dehoog_sm_col_entries(d->n) * (d->dtype == NL_F64 ? 16 : 8);
  for (int th = 128; th >= 32; th /= 2)
    if (per_thread * th <= 72 * 1024) return th;
  return 0;
}
size_t sm_smem(const nl_desc* d) {
  return (size_t)dehoog_sm_col_entries(d->n) * (d->dtype == NL_F64 ? 16 : 8) * sm_threads(d);
}
template <typename R, bool BWD, int THREADS>
int sm_blocks(size_t smem) {
  int v = 0;
  cudaFuncSetAttribute(dehoog_qd_sm_kernel<R, BWD, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dehoog_qd_sm_kernel<R, BWD, THREADS>, THREADS, smem);
  return (v > 0 ? v : 1) * device_sm_count();
}
template <typename R>
int64_t sm_cap(int th, bool bwd, size_t smem) {
  if (th == 128) return bwd ? sm_blocks<R, true, 128>(smem) : sm_blocks<R, false, 128>(smem);
  if (th == 64) return bwd ? sm_blocks<R, true, 64>(smem) : sm_blocks<R, false, 64>(smem);
  return bwd ? sm_blocks<R, true, 32>(smem) : sm_blocks<R, false, 32>(smem);
}
// launched CTAs (persistent, grid-stride): the table is one slot per launched thread
int64_t qd_grid(const nl_desc* d, bool bwd) {
  const int64_t items = (int64_t)d->B * d->T * d->D;
  const int th = sm_threads(d);
  if (th == 0) {
    const int64_t need = (items + kQdThreads - 1) / kQdThreads;
    return std::min<int64_t>(need, (int64_t)device_sm_count() * 8);
  }
  const int64_t need = (items + th - 1) / th;
  const int64_t cap = d->dtype == NL_F64 ? sm_cap<double>(th, bwd, sm_smem(d)) : sm_cap<float>(th, bwd, sm_smem(d));
  return std::min<int64_t>(need, cap);
}

template <typename R, bool BWD, int THREADS>
void sm_launch(const QdArgs<R>& A, int64_t blocks, size_t smem, cudaStream_t st) {
  dehoog_qd_sm_kernel<R, BWD, THREADS><<<(unsigned)blocks, THREADS, smem, st>>>(A);
}

template <typename R>
cudaError_t launch_t(const nl_desc* d, bool bwd, const void* t, const void* Tper, const void* F, const void* gx,
                     void* x, void* dF, void* scratch, cudaStream_t st) {
  const int64_t blocks = qd_grid(d, bwd);
  if (blocks == 0) return cudaSuccess;
  QdArgs<R> A{};
  A.P.family = NL_DEHOOG;
  A.P.n = d->n;
  A.P.head = d->head;
  A.P.alpha = (R)d->alpha;
  A.P.log_tol = (R)d->log_tol;
  A.P.scale = (R)d->scale;
  A.B = d->B; A.T = d->T; A.D = d->D;
  A.per_row = d->t_mode == NL_T_PER_ROW ? 1 : 0;
  A.t = (const R*)t; A.Tper = (const R*)Tper; A.F = (const Cx<R>*)F; A.gx = (const R*)gx; A.x = (R*)x; A.dF = (Cx<R>*)dF;
  A.scratch = (Cx<R>*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  const int th = sm_threads(d);
  if (th == 0) {
    A.stride = blocks * kQdThreads;
    if (bwd) dehoog_qd_kernel<R, true><<<(unsigned)blocks, kQdThreads, 0, st>>>(A);
    else dehoog_qd_kernel<R, false><<<(unsigned)blocks, kQdThreads, 0, st>>>(A);
    return cudaGetLastError();
  }
  A.stride = blocks * th;
  const size_t smem = sm_smem(d);
  if (th == 128) {
    if (bwd) sm_launch<R, true, 128>(A, blocks, smem, st);
    else sm_launch<R, false, 128>(A, blocks, smem, st);
  } else if (th == 64) {
    if (bwd) sm_launch<R, true, 64>(A, blocks, smem, st);
    else sm_launch<R, false, 64>(A, blocks, smem, st);
  } else {
    if (bwd) sm_launch<R, true, 32>(A, blocks, smem, st);
    else sm_launch<R, false, 32>(A, blocks, smem, st);
  }
  return cudaGetLastError();
}
}  // namespace

size_t dehoog_qd_scratch_bytes(const nl_desc* d, bool bwd) {
  const size_t cs = d->dtype == NL_F64 ? 16 : 8;
  const int th = sm_threads(d);
  if (th == 0)
    return (size_t)dehoog_scratch_entries(d->n, bwd) * (size_t)qd_grid(d, bwd) * kQdThreads * cs + 256;
  if (!bwd) return 256;
  return (size_t)dehoog_sm_tab_entries(d->n) * (size_t)qd_grid(d, true) * th * cs + 256;
}

cudaError_t launch_dehoog_qd(const nl_desc* d, bool bwd, const void* t, const void* F, const void* gx, void* x,
                             void* dF, void* scratch, cudaStream_t st, const void* Tper) {
  return d->dtype == NL_F64 ? launch_t<double>(d, bwd, t, Tper, F, gx, x, dF, scratch, st)
                            : launch_t<float>(d, bwd, t, Tper, F, gx, x, dF, scratch, st);
}

}  // namespace nl
