// Stand-alone (split-path) kernels: query points + sphere projection, sphere maps, and the
// line-integrate reductions (forward + backward) on explicit F tensors.  These serve the
// class-level ILT API and arbitrary user laplace_rep_func modules; the fused kernels in
// nl_fused_simt.cu / nl_fused_tc.cu are the hot path.
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
                                      R* __restrict__ gb, Cx<R>* scratch, int64_t scratch_stride) {
  const int n = P.n;
  const int64_t items = (int64_t)B * T * D;
  FView<R> F{fa, fb, layout, NL_HEAD_SPHERE};
  // grid-stride so the DeHoog scratch (one slot per launched thread) stays bounded
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = item / D;
    int64_t tidx = (t_mode == NL_T_PER_ROW) ? row : (row % T);
    TimeCtx<R> c = make_time_ctx(P, t[tidx], Tper, tidx);
    const int64_t off0 = item * n;
    if (P.family != NL_DEHOOG) {
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
    } else {
      const int M = (n - 1) / 2;
      Scratch<R> scr{scratch, scratch_stride, (int64_t)blockIdx.x * blockDim.x + threadIdx.x};
      R sn, cs;
      m_sincos((R(kPi) * c.t) / c.T, &sn, &cs);  // z = exp(i*pi*t/T), inverse_laplace.py:842
      Cx<R> z{cs, sn};
      auto a = [&](int k) {
        HeadOut<R> h = F.load(off0 + k);
        return Cx<R>{h.fre, h.fim};
      };
      if (!BWD) {
        Cx<R> w = dehoog_forward<R>(M, z, a, scr);
        x[item] = c.pref * w.re;
      } else {
        R g = gx[item] * c.pref;
        auto emit = [&](int k, Cx<R> dw) {
          HeadOut<R> h = F.load(off0 + k);
          store_grad(layout, ga, gb, off0 + k, h, g * dw.re, -g * dw.im);
        };
        dehoog_forward_backward<R>(M, z, a, emit, scr);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------
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
  Cx<R>* scratch = nullptr;
  int64_t stride = 0;
  if (d->family == NL_DEHOOG) {
    // bounded grid; scratch allocated stream-ordered (the split path is not the hot path)
    int64_t max_blocks = 148 * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    stride = blocks * threads;
    size_t bytes = (size_t)dehoog_scratch_entries(d->n, bwd) * stride * sizeof(Cx<R>);
    cudaError_t e = cudaMallocAsync((void**)&scratch, bytes, st);
    if (e != cudaSuccess) return e;
  }
  if (bwd)
    line_integrate_kernel<R, true><<<(unsigned)blocks, threads, 0, st>>>(
        P, d->B, d->T, d->D, d->t_mode, layout, (const R*)fa, (const R*)fb, (const R*)t, (const R*)Tper,
        (const R*)gx, nullptr, (R*)ga, (R*)gb, scratch, stride);
  else
    line_integrate_kernel<R, false><<<(unsigned)blocks, threads, 0, st>>>(
        P, d->B, d->T, d->D, d->t_mode, layout, (const R*)fa, (const R*)fb, (const R*)t, (const R*)Tper, nullptr,
        (R*)x, nullptr, nullptr, scratch, stride);
  cudaError_t e = cudaGetLastError();
  if (scratch) {
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = e2;
  }
  return e;
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
