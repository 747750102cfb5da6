// Stand-alone (split-path) kernels: query points + sphere projection, sphere maps, and the
// line-integrate reductions (forward + backward) on explicit F tensors.  These serve the
// class-level ILT API and arbitrary user laplace_rep_func modules; the fused kernels in
// nl_fused_simt.cu / nl_fused_tc.cu are the hot path.
#include <algorithm>

#include "nl_dehoog.cuh"
#include "nl_kernels.h"

namespace nl {

// ---------------------------------------------------------------------------------------
// compute_s (+ complex_to_spherical_riemann): one thread per (row, k).
// ---------------------------------------------------------------------------------------
template <typename R>
__global__ void query_points_kernel(IltParams<R> P, int64_t rows, const R* __restrict__ t,
                                    const R* __restrict__ Tper, R* __restrict__ s_out,
                                    R* __restrict__ feat_out) {
  const int n = P.n;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  int64_t row = idx / n;
  int k = (int)(idx - row * n);
  TimeCtx<R> c = make_time_ctx(P, t[row], Tper, row);
  Cx<R> s = s_point(P, c, k);
  if (s_out) {
    s_out[2 * idx] = s.re;
    s_out[2 * idx + 1] = s.im;
  }
  if (feat_out) {
    R f0, f1;
    s_features(P, s, &f0, &f1);
    feat_out[row * 2 * n + k] = f0;
    feat_out[row * 2 * n + n + k] = f1;
  }
}

template <typename R>
__global__ void complex_to_sphere_kernel(int64_t count, const R* __restrict__ sr, const R* __restrict__ si,
                                         R* __restrict__ theta, R* __restrict__ phi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  IltParams<R> P{};
  P.head = NL_HEAD_SPHERE;
  R f0, f1;
  s_features(P, Cx<R>{sr[i], si[i]}, &f0, &f1);
  theta[i] = f0;
  phi[i] = f1;
}

template <typename R>
__global__ void sphere_to_complex_kernel(int64_t count, const R* __restrict__ theta, const R* __restrict__ phi,
                                         R* __restrict__ sr, R* __restrict__ si) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  // transformations.py:50-53
  R r = m_tan(phi[i] / R(2) + R(kPi) / R(4));
  R sn, cs;
  m_sincos(theta[i], &sn, &cs);
  sr[i] = r * cs;
  si[i] = r * sn;
}

// ---------------------------------------------------------------------------------------
// line integrate: one thread per (b, t, d) item.
// ---------------------------------------------------------------------------------------
template <typename R>
struct FView {
  const R* a;
  const R* b;
  int layout;
  int head;  // NL_HEAD_SPHERE when layout == SPHERE_PLANAR
  __device__ __forceinline__ HeadOut<R> load(int64_t off) const {
    if (layout == NL_F_COMPLEX_INTERLEAVED) return head_forward<R>(NL_HEAD_RAW, a[2 * off], a[2 * off + 1]);
    if (layout == NL_F_COMPLEX_PLANAR) return head_forward<R>(NL_HEAD_RAW, a[off], b[off]);
    // planar sphere angles: F = tan(phi/2+pi/4) e^{i theta}; no tanh head here (user module did it)
    HeadOut<R> h;
    R th = a[off], ph = b[off];
    h.r = m_tan(ph / R(2) + R(kPi) / R(4));
    m_sincos(th, &h.sn, &h.cs);
    h.fre = h.r * h.cs;
    h.fim = h.r * h.sn;
    h.ta = h.tb = R(0);
    return h;
  }
};

template <typename R>
__device__ __forceinline__ void store_grad(int layout, R* ga, R* gb, int64_t off, const HeadOut<R>& h, R gre, R gim) {
  if (layout == NL_F_COMPLEX_INTERLEAVED) {
    ga[2 * off] = gre;
    ga[2 * off + 1] = gim;
  } else if (layout == NL_F_COMPLEX_PLANAR) {
    ga[off] = gre;
    gb[off] = gim;
  } else {
    ga[off] = h.r * (gim * h.cs - gre * h.sn);                                // d/dtheta
    gb[off] = (gre * h.cs + gim * h.sn) * (R(1) + h.r * h.r) * R(0.5);       // d/dphi
  }
}

template <typename R, bool BWD>
__global__ void line_integrate_kernel(IltParams<R> P, int B, int T, int D, int t_mode, int layout,
                                      const R* __restrict__ fa, const R* __restrict__ fb,
                                      const R* __restrict__ t, const R* __restrict__ Tper,
                                      const R* __restrict__ gx, R* __restrict__ x, R* __restrict__ ga,
                                      R* __restrict__ gb) {
  const int n = P.n;
  const int64_t items = (int64_t)B * T * D;
  FView<R> F{fa, fb, layout, NL_HEAD_SPHERE};
  // (De Hoog does not come here: launch_line_integrate hands it to the recurrence kernels of nl_dehoog_qd.cu)
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = item / D;
    int64_t tidx = (t_mode == NL_T_PER_ROW) ? row : (row % T);
    TimeCtx<R> c = make_time_ctx(P, t[tidx], Tper, tidx);
    const int64_t off0 = item * n;
    {
      R acc = R(0);
      R g = BWD ? gx[item] * c.pref : R(0);
      for (int k = 0; k < n; ++k) {
        HeadOut<R> h = F.load(off0 + k);
        Cx<R> w = reduce_coef(P, c, k);
        if (!BWD) {
          acc += h.fre * w.re - h.fim * w.im;
        } else {
          store_grad(layout, ga, gb, off0 + k, h, g * w.re, -g * w.im);
        }
      }
      if (!BWD) x[item] = c.pref * acc;
    }
  }
}

// De Hoog: the recurrence runs in the full-occupancy kernels of nl_dehoog_qd.cu, which work on [d][k][t][b]
// complex planes.  to_planes: F in the caller's layout -> plane; from_planes: dL/dF plane (already g Re dw, -g Im dw)
// -> gradient in the caller's layout (chain rule of the sphere angles included).  One thread per (d, k, t, b),
// b fastest: plane accesses coalesced, caller-layout accesses strided by D*n.
template <typename R, bool BACK>
__global__ void dehoog_planes_kernel(int B, int T, int D, int n, int layout, const R* __restrict__ fa,
                                     const R* __restrict__ fb, Cx<R>* __restrict__ plane_f,
                                     const Cx<R>* __restrict__ plane_df, R* __restrict__ ga, R* __restrict__ gb) {
  const int64_t tb_count = (int64_t)T * B;
  const int64_t total = tb_count * D * n;
  FView<R> F{fa, fb, layout, NL_HEAD_SPHERE};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dk = i / tb_count, tb = i - dk * tb_count;
    const int d = (int)(dk / n), k = (int)(dk - (int64_t)d * n);
    const int ti = (int)(tb / B), b = (int)(tb - (int64_t)ti * B);
    const int64_t off = (((int64_t)b * T + ti) * D + d) * n + k;
    const HeadOut<R> h = F.load(off);
    if (!BACK) {
      plane_f[i] = Cx<R>{h.fre, h.fim};
    } else {
      const Cx<R> df = plane_df[i];
      store_grad(layout, ga, gb, off, h, df.re, df.im);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------
// Caller-owned scratch for the entry points that have no workspace argument (nl_set_scratch); the De Hoog line
// integral needs it (planes + recurrence table) and fails with NL_ERR_WORKSPACE without it.
static thread_local void* g_scratch = nullptr;
static thread_local size_t g_scratch_bytes = 0;
void set_thread_scratch(void* ptr, size_t bytes) {
  g_scratch = ptr;
  g_scratch_bytes = ptr ? bytes : 0;
}
size_t thread_scratch_bytes() { return g_scratch_bytes; }
size_t line_integrate_scratch_bytes(const nl_desc* d, bool bwd) {
  if (d->family != NL_DEHOOG) return 0;
  const size_t rs = d->dtype == NL_F64 ? 16 : 8;
  const size_t plane_bytes = ((size_t)d->B * d->T * d->D * d->n * rs + 255) / 256 * 256;
  return plane_bytes * (bwd ? 2 : 1) + dehoog_qd_scratch_bytes(d, bwd) + 512;
}

template <typename R>
static IltParams<R> make_params(const nl_desc* d, const void* beta, const void* eta) {
  IltParams<R> P;
  P.family = d->family;
  P.n = d->n;
  P.head = d->head;
  P.alpha = (R)d->alpha;
  P.log_tol = (R)d->log_tol;
  P.scale = (R)d->scale;
  P.beta = (const R*)beta;
  P.eta = (const R*)eta;
  return P;
}

template <typename R>
cudaError_t launch_query_points(const nl_desc* d, int64_t rows, const void* t, const void* Tper, const void* beta,
                                void* s_out, void* feat_out, cudaStream_t st) {
  IltParams<R> P = make_params<R>(d, beta, nullptr);
  int64_t total = rows * d->n;
  if (total == 0) return cudaSuccess;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  query_points_kernel<R><<<(unsigned)blocks, threads, 0, st>>>(P, rows, (const R*)t, (const R*)Tper, (R*)s_out,
                                                             (R*)feat_out);
  return cudaGetLastError();
}

template <typename R>
cudaError_t launch_sphere(int which, int64_t count, const void* a, const void* b, void* c, void* e, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  int threads = 256;
  int64_t blocks = (count + threads - 1) / threads;
  if (which == 0)
    complex_to_sphere_kernel<R><<<(unsigned)blocks, threads, 0, st>>>(count, (const R*)a, (const R*)b, (R*)c, (R*)e);
  else
    sphere_to_complex_kernel<R><<<(unsigned)blocks, threads, 0, st>>>(count, (const R*)a, (const R*)b, (R*)c, (R*)e);
  return cudaGetLastError();
}

template <typename R>
cudaError_t launch_line_integrate(const nl_desc* d, int layout, bool bwd, const void* fa, const void* fb,
                                  const void* t, const void* Tper, const void* eta, const void* gx, void* x,
                                  void* ga, void* gb, cudaStream_t st) {
  IltParams<R> P = make_params<R>(d, nullptr, eta);
  int64_t items = (int64_t)d->B * d->T * d->D;
  if (items == 0) return cudaSuccess;
  int threads = 128;
  int64_t blocks = (items + threads - 1) / threads;
  if (d->family == NL_DEHOOG) {
    // planes + recurrence scratch come from the caller (nl_set_scratch): the library never allocates
    const size_t plane_bytes = ((size_t)items * d->n * sizeof(Cx<R>) + 255) / 256 * 256;
    const size_t need = line_integrate_scratch_bytes(d, bwd);
    if (!(g_scratch && g_scratch_bytes >= need)) return cudaErrorInvalidValue;  // nl_api.cu checks first
    unsigned char* buf = (unsigned char*)(((uintptr_t)g_scratch + 255) & ~(uintptr_t)255);
    cudaError_t e = cudaSuccess;
    Cx<R>* pf = (Cx<R>*)buf;
    Cx<R>* pdf = bwd ? (Cx<R>*)(buf + plane_bytes) : nullptr;
    void* scr = buf + plane_bytes * (bwd ? 2 : 1);
    const int64_t total = items * d->n;
    const int cb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)device_sm_count() * 16);
    dehoog_planes_kernel<R, false><<<cb, 256, 0, st>>>(d->B, d->T, d->D, d->n, layout, (const R*)fa, (const R*)fb, pf,
                                                      nullptr, nullptr, nullptr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = launch_dehoog_qd(d, bwd, t, pf, gx, x, pdf, scr, st, Tper);
    if (e == cudaSuccess && bwd) {
      dehoog_planes_kernel<R, true><<<cb, 256, 0, st>>>(d->B, d->T, d->D, d->n, layout, (const R*)fa, (const R*)fb,
                                                       nullptr, pdf, (R*)ga, (R*)gb);
      e = cudaGetLastError();
    }
    return e;
  }
  if (bwd)
    line_integrate_kernel<R, true><<<(unsigned)blocks, threads, 0, st>>>(
        P, d->B, d->T, d->D, d->t_mode, layout, (const R*)fa, (const R*)fb, (const R*)t, (const R*)Tper,
        (const R*)gx, nullptr, (R*)ga, (R*)gb);
  else
    line_integrate_kernel<R, false><<<(unsigned)blocks, threads, 0, st>>>(
        P, d->B, d->T, d->D, d->t_mode, layout, (const R*)fa, (const R*)fb, (const R*)t, (const R*)Tper, nullptr,
        (R*)x, nullptr, nullptr);
  return cudaGetLastError();
}

template cudaError_t launch_query_points<float>(const nl_desc*, int64_t, const void*, const void*, const void*,
                                                void*, void*, cudaStream_t);
template cudaError_t launch_query_points<double>(const nl_desc*, int64_t, const void*, const void*, const void*,
                                                 void*, void*, cudaStream_t);
template cudaError_t launch_sphere<float>(int, int64_t, const void*, const void*, void*, void*, cudaStream_t);
template cudaError_t launch_sphere<double>(int, int64_t, const void*, const void*, void*, void*, cudaStream_t);
template cudaError_t launch_line_integrate<float>(const nl_desc*, int, bool, const void*, const void*, const void*,
                                                  const void*, const void*, const void*, void*, void*, void*,
                                                  cudaStream_t);
template cudaError_t launch_line_integrate<double>(const nl_desc*, int, bool, const void*, const void*,
                                                   const void*, const void*, const void*, const void*, void*, void*,
                                                   void*, cudaStream_t);

}  // namespace nl
