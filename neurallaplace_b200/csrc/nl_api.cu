// C ABI of the library (see include/nl_b200.h for the contract and the reference lines each
// entry point replaces).  Thin: validate, pick the kernel family, launch on the caller's stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nl_kernels.h"

namespace {
thread_local char g_cuda_err[256] = "";
thread_local int g_launches = 0;

int cuda_fail(cudaError_t e) {
  std::snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return NL_ERR_CUDA;
}

int check_desc(const nl_desc* d) {
  if (!d) return NL_ERR_NULL;
  if (d->dtype != NL_F32 && d->dtype != NL_F64) return NL_ERR_ENUM;
  if (d->family < NL_FOURIER || d->family > NL_DEHOOG) return NL_ERR_ENUM;
  if (d->t_mode != NL_T_SHARED && d->t_mode != NL_T_PER_ROW) return NL_ERR_ENUM;
  if (d->head != NL_HEAD_SPHERE && d->head != NL_HEAD_RAW) return NL_ERR_ENUM;
  if (d->impl < 0 || d->impl > 2) return NL_ERR_ENUM;
  return NL_OK;
}

int check_fused(const nl_desc* d) {
  int s = check_desc(d);
  if (s) return s;
  if (d->B < 0 || d->T < 0 || d->K < 1 || d->H < 1 || d->D < 1 || d->n < 1) return NL_ERR_SHAPE;
  return NL_OK;
}

// 1 = SIMT, 2 = tcgen05, 0 = neither covers the shape
bool tc_any(const nl_desc* d, int pass) {
  return nl::tc2_plan(d, pass).ok || nl::tc_plan(d, pass).ok || nl::tc_pr_plan(d, pass).ok ||
         nl::tc_dh_plan(d, pass).ok;
}

int pick_impl(const nl_desc* d, int pass) {
  if (d->impl == 1) return nl::simt_plan(d, pass).ok ? 1 : 0;
  if (d->impl == 2) return tc_any(d, pass) ? 2 : 0;
  if (tc_any(d, pass)) return 2;
  if (nl::simt_plan(d, pass).ok) return 1;
  return 0;
}
}  // namespace

extern "C" {

int nl_abi_version(void) { return NL_B200_ABI_VERSION; }

const char* nl_strerror(int status) {
  switch (status) {
    case NL_OK: return "ok";
    case NL_ERR_NULL: return "required pointer is NULL";
    case NL_ERR_SHAPE: return "invalid dimension";
    case NL_ERR_UNSUPPORTED: return "shape not covered by the fused kernels (use the split path)";
    case NL_ERR_WORKSPACE: return "workspace / scratch missing or too small";
    case NL_ERR_CUDA: return "CUDA runtime error";
    case NL_ERR_ENUM: return "invalid enum value in nl_desc";
    default: return "unknown nl_status";
  }
}

const char* nl_last_cuda_error(void) { return g_cuda_err; }
int nl_last_launch_count(void) { return g_launches; }

size_t nl_workspace_bytes(const nl_desc* d, int pass) {
  if (check_fused(d)) return 0;
  size_t a = 0, b = 0;
  nl::SimtPlan sp = nl::simt_plan(d, pass);
  if (sp.ok) a = sp.ws_bytes;
  nl::TcPlan tp = nl::tc_plan(d, pass);
  if (tp.ok) b = tp.ws_bytes;
  nl::TcPlan tp2 = nl::tc2_plan(d, pass);
  if (tp2.ok && tp2.ws_bytes > b) b = tp2.ws_bytes;
  nl::TcPlan tpr = nl::tc_pr_plan(d, pass);
  if (tpr.ok && tpr.ws_bytes > b) b = tpr.ws_bytes;
  nl::TcPlan tdh = nl::tc_dh_plan(d, pass);
  if (tdh.ok && tdh.ws_bytes > b) b = tdh.ws_bytes;
  return (a > b ? a : b) + 256;
}

int nl_selected_impl(const nl_desc* d, int pass) {
  if (check_fused(d)) return 0;
  return pick_impl(d, pass);
}

static size_t saved_bytes(const nl_desc* d) {
  if (check_fused(d)) return 0;
  if (pick_impl(d, 0) != 2 || pick_impl(d, 1) != 2) return 0;
  if (d->t_mode == NL_T_PER_ROW) return nl::tc_pr_saved_bytes(d);
  if (d->family == NL_DEHOOG) return nl::tc_dh_saved_bytes(d);
  return nl::tc_saved_bytes(d);
}

static int reconstruct(const nl_desc* d, bool bwd, const void* p, const void* t, const void* const* weights,
                       const void* beta, const void* eta, void* x, const void* gx, void* const* grads, void* ws,
                       size_t ws_bytes, void* stream, void* saved = nullptr, size_t saved_size = 0) {
  g_launches = 0;
  int s = check_fused(d);
  if (s) return s;
  const bool empty = (int64_t)d->B * d->T == 0;  // empty tensors legitimately have NULL data pointers
  if (!weights) return NL_ERR_NULL;
  if (!empty && (!p || !t)) return NL_ERR_NULL;
  for (int i = 0; i < 6; ++i)
    if (!weights[i]) return NL_ERR_NULL;
  if (d->family == NL_AW && (!beta || !eta)) return NL_ERR_NULL;
  if (bwd) {
    if (!grads || (!empty && !gx)) return NL_ERR_NULL;
    for (int i = 0; i < 6; ++i)
      if (!grads[i]) return NL_ERR_NULL;
    if (d->B > 0 && !grads[6]) return NL_ERR_NULL;
  } else if (!empty && !x) {
    return NL_ERR_NULL;
  }
  const int pass = bwd ? 1 : 0;
  int impl = pick_impl(d, pass);
  // per-row time grids / De Hoog: the tensor-core backward exists only in its saved form
  // (nl_reconstruct_backward_saved)
  if (impl == 2 && bwd && !saved && (d->t_mode == NL_T_PER_ROW || d->family == NL_DEHOOG))
    impl = (d->impl != 2 && nl::simt_plan(d, pass).ok) ? 1 : 0;
  if (impl == 0) return NL_ERR_UNSUPPORTED;
  if (saved) {
    const size_t sb = saved_bytes(d);
    if (sb == 0) return NL_ERR_UNSUPPORTED;
    if (saved_size < sb) return NL_ERR_WORKSPACE;
  }
  const size_t need = nl_workspace_bytes(d, pass);
  if (need > 0 && (!ws || ws_bytes < need)) return NL_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (impl == 2 && d->t_mode == NL_T_PER_ROW) {
    e = nl::launch_fused_tc_per_row(d, bwd, p, t, weights, beta, eta, x, gx, grads, ws, st, &g_launches, saved);
  } else if (impl == 2 && d->family == NL_DEHOOG) {
    e = nl::launch_fused_tc_dehoog(d, bwd, p, t, weights, x, gx, grads, ws, st, &g_launches, saved);
  } else if (impl == 2) {
    // Two tensor-core variants exist (nl_fused_tc.cu: one stream per CTA, nl_fused_tc2.cu: two streams + a
    // dedicated MMA-issuer warp).  Measured on B200: backward -> one stream; forward -> one stream at 2 CTAs/SM
    // while W3 is small enough for that, else two streams.  NL_B200_TC_STREAMS=1|2 forces one (profiling).
    static const int force = [] { const char* v = getenv("NL_B200_TC_STREAMS"); return v ? atoi(v) : 0; }();
    const nl::TcPlan p1 = nl::tc_plan(d, pass);
    const bool ok1 = p1.ok, ok2 = nl::tc2_plan(d, pass).ok;
    // forward: one-stream kernel when two of its CTAs co-reside on an SM, else the two-stream kernel
    // (measured: with every W3 chunk resident the two-stream kernel beats two co-resident one-stream CTAs that
    // stream W3 through a ring -- d_obs 4 x 33 terms: 2.12 vs 2.39 ms; beyond its shared-memory limit the ring wins)
    bool use2 = bwd ? false : !(ok1 && p1.co_resident && !(p1.ring && ok2));
    if (force == 1) use2 = false;
    if (force == 2) use2 = true;
    if (use2 && !ok2) use2 = false;
    if (!use2 && !ok1) use2 = true;
    if (saved) use2 = false;  // the activation stash is written / read by the one-stream kernels
    if (use2)
      e = nl::launch_fused_tc2(d, bwd, p, t, weights, beta, eta, x, gx, grads, ws, st, &g_launches);
    else
      e = nl::launch_fused_tc(d, bwd, p, t, weights, beta, eta, x, gx, grads, ws, st, &g_launches, empty ? nullptr : saved);
  } else if (d->dtype == NL_F64) {
    e = nl::launch_fused_simt<double>(d, bwd, p, t, weights, beta, eta, x, gx, grads, ws, st, &g_launches);
  } else {
    e = nl::launch_fused_simt<float>(d, bwd, p, t, weights, beta, eta, x, gx, grads, ws, st, &g_launches);
  }
  if (e != cudaSuccess) return cuda_fail(e);
  return NL_OK;
}

int nl_reconstruct_forward(const nl_desc* d, const void* p, const void* t, const void* const* weights,
                           const void* beta, const void* eta, void* x, void* ws, size_t ws_bytes, void* stream) {
  return reconstruct(d, false, p, t, weights, beta, eta, x, nullptr, nullptr, ws, ws_bytes, stream);
}

int nl_reconstruct_backward(const nl_desc* d, const void* p, const void* t, const void* const* weights,
                            const void* beta, const void* eta, const void* grad_x, void* const* grads, void* ws,
                            size_t ws_bytes, void* stream) {
  return reconstruct(d, true, p, t, weights, beta, eta, nullptr, grad_x, grads, ws, ws_bytes, stream);
}

size_t nl_saved_bytes(const nl_desc* d) { return saved_bytes(d); }

int nl_reconstruct_forward_save(const nl_desc* d, const void* p, const void* t, const void* const* weights,
                                const void* beta, const void* eta, void* x, void* saved, size_t saved_size, void* ws,
                                size_t ws_bytes, void* stream) {
  if (!saved) return NL_ERR_NULL;
  return reconstruct(d, false, p, t, weights, beta, eta, x, nullptr, nullptr, ws, ws_bytes, stream, saved, saved_size);
}

int nl_reconstruct_backward_saved(const nl_desc* d, const void* p, const void* t, const void* const* weights,
                                  const void* beta, const void* eta, const void* grad_x, const void* saved,
                                  size_t saved_size, void* const* grads, void* ws, size_t ws_bytes, void* stream) {
  if (!saved) return NL_ERR_NULL;
  return reconstruct(d, true, p, t, weights, beta, eta, nullptr, grad_x, grads, ws, ws_bytes, stream,
                     const_cast<void*>(saved), saved_size);
}

int nl_query_points(const nl_desc* d, int64_t rows, const void* t, const void* Tper, const void* beta, void* s_out,
                    void* feat_out, void* stream) {
  int s = check_desc(d);
  if (s) return s;
  if (rows < 0 || d->n < 1) return NL_ERR_SHAPE;
  if (rows == 0) return NL_OK;
  if (!t) return NL_ERR_NULL;
  if (d->family == NL_AW && !beta) return NL_ERR_NULL;
  cudaError_t e = d->dtype == NL_F64
                      ? nl::launch_query_points<double>(d, rows, t, Tper, beta, s_out, feat_out, (cudaStream_t)stream)
                      : nl::launch_query_points<float>(d, rows, t, Tper, beta, s_out, feat_out, (cudaStream_t)stream);
  return e == cudaSuccess ? NL_OK : cuda_fail(e);
}

static int line_integrate(const nl_desc* d, int layout, bool bwd, const void* fa, const void* fb, const void* t,
                          const void* Tper, const void* eta, const void* gx, void* x, void* ga, void* gb,
                          void* stream) {
  int s = check_desc(d);
  if (s) return s;
  if (layout < NL_F_SPHERE_PLANAR || layout > NL_F_COMPLEX_INTERLEAVED) return NL_ERR_ENUM;
  if (d->B < 0 || d->T < 0 || d->D < 1 || d->n < 1) return NL_ERR_SHAPE;
  if (d->family == NL_DEHOOG && (d->n < 3 || d->n % 2 == 0)) return NL_ERR_SHAPE;
  const bool planar = layout != NL_F_COMPLEX_INTERLEAVED;
  if ((int64_t)d->B * d->T == 0) return NL_OK;
  if (!fa || !t || (planar && !fb)) return NL_ERR_NULL;
  if (d->family == NL_AW && !eta) return NL_ERR_NULL;
  if (bwd ? (!gx || !ga || (planar && !gb)) : !x) return NL_ERR_NULL;
  // De Hoog needs caller-owned scratch (nl_set_scratch): the library never allocates
  if (nl::thread_scratch_bytes() < nl::line_integrate_scratch_bytes(d, bwd)) return NL_ERR_WORKSPACE;
  cudaError_t e =
      d->dtype == NL_F64
          ? nl::launch_line_integrate<double>(d, layout, bwd, fa, fb, t, Tper, eta, gx, x, ga, gb, (cudaStream_t)stream)
          : nl::launch_line_integrate<float>(d, layout, bwd, fa, fb, t, Tper, eta, gx, x, ga, gb, (cudaStream_t)stream);
  return e == cudaSuccess ? NL_OK : cuda_fail(e);
}

int nl_line_integrate_forward(const nl_desc* d, int f_layout, const void* fa, const void* fb, const void* t,
                              const void* Tper, const void* eta, void* x, void* stream) {
  return line_integrate(d, f_layout, false, fa, fb, t, Tper, eta, nullptr, x, nullptr, nullptr, stream);
}

int nl_line_integrate_backward(const nl_desc* d, int f_layout, const void* fa, const void* fb, const void* t,
                               const void* Tper, const void* eta, const void* grad_x, void* grad_fa, void* grad_fb,
                               void* stream) {
  return line_integrate(d, f_layout, true, fa, fb, t, Tper, eta, grad_x, nullptr, grad_fa, grad_fb, stream);
}

static int sphere(int which, int dtype, int64_t count, const void* a, const void* b, void* c, void* e, void* stream) {
  if (dtype != NL_F32 && dtype != NL_F64) return NL_ERR_ENUM;
  if (count < 0) return NL_ERR_SHAPE;
  if (count > 0 && (!a || !b || !c || !e)) return NL_ERR_NULL;
  cudaError_t err = dtype == NL_F64 ? nl::launch_sphere<double>(which, count, a, b, c, e, (cudaStream_t)stream)
                                    : nl::launch_sphere<float>(which, count, a, b, c, e, (cudaStream_t)stream);
  return err == cudaSuccess ? NL_OK : cuda_fail(err);
}

int nl_complex_to_sphere(int dtype, int64_t count, const void* s_re, const void* s_im, void* theta, void* phi,
                         void* stream) {
  return sphere(0, dtype, count, s_re, s_im, theta, phi, stream);
}

int nl_sphere_to_complex(int dtype, int64_t count, const void* theta, const void* phi, void* s_re, void* s_im,
                         void* stream) {
  return sphere(1, dtype, count, theta, phi, s_re, s_im, stream);
}

size_t nl_line_integrate_scratch_bytes(const nl_desc* d, int pass) {
  if (check_desc(d) || d->B < 0 || d->T < 0 || d->D < 1 || d->n < 1) return 0;
  if ((int64_t)d->B * d->T == 0) return 0;
  return nl::line_integrate_scratch_bytes(d, pass != 0);
}
void nl_set_scratch(void* ptr, size_t bytes) { nl::set_thread_scratch(ptr, bytes); }

static int dehoog_fixed(int dtype, int64_t rows, int n, int64_t T, bool bwd, const void* F, const void* t,
                        double contour_T, double gamma, const void* gx, void* x, void* dF, void* scratch,
                        size_t scratch_bytes, void* stream) {
  if (dtype != NL_F32 && dtype != NL_F64) return NL_ERR_ENUM;
  if (rows < 0 || T < 0 || n < 3 || n % 2 == 0) return NL_ERR_SHAPE;
  if (!(contour_T > 0.0)) return NL_ERR_SHAPE;
  if (rows == 0 || T == 0) return NL_OK;
  if (!F || !t || !scratch || (bwd ? (!gx || !dF) : !x)) return NL_ERR_NULL;
  if (scratch_bytes < nl::dehoog_fixed_scratch_bytes(dtype, rows, n, T, bwd)) return NL_ERR_WORKSPACE;
  g_launches = bwd ? 3 : 2;
  cudaError_t e = nl::launch_dehoog_fixed(dtype, rows, n, T, bwd, F, t, contour_T, gamma, gx, x, dF, scratch,
                                          (cudaStream_t)stream);
  return e == cudaSuccess ? NL_OK : cuda_fail(e);
}

size_t nl_dehoog_fixed_scratch_bytes(int dtype, int64_t rows, int n, int64_t T, int pass) {
  if ((dtype != NL_F32 && dtype != NL_F64) || rows <= 0 || T <= 0 || n < 3 || n % 2 == 0) return 0;
  return nl::dehoog_fixed_scratch_bytes(dtype, rows, n, T, pass != 0);
}

int nl_dehoog_fixed_forward(int dtype, int64_t rows, int n, int64_t T, const void* F, const void* t,
                            double contour_T, double gamma, void* x, void* scratch, size_t scratch_bytes,
                            void* stream) {
  return dehoog_fixed(dtype, rows, n, T, false, F, t, contour_T, gamma, nullptr, x, nullptr, scratch, scratch_bytes,
                      stream);
}

int nl_dehoog_fixed_backward(int dtype, int64_t rows, int n, int64_t T, const void* F, const void* t,
                             double contour_T, double gamma, const void* grad_x, void* grad_F, void* scratch,
                             size_t scratch_bytes, void* stream) {
  return dehoog_fixed(dtype, rows, n, T, true, F, t, contour_T, gamma, grad_x, nullptr, grad_F, scratch,
                      scratch_bytes, stream);
}

size_t nl_clip_adam_scratch_bytes(void) { return nl::clip_adam_scratch_bytes(); }

int nl_clip_adam_step(int dtype, int64_t count, void* params, void* grads, void* exp_avg, void* exp_avg_sq, double lr,
                      double beta1, double beta2, double eps, int64_t step, double max_norm, int write_clipped_grads,
                      void* scratch, size_t scratch_bytes, void* stream) {
  if (dtype != NL_F32 && dtype != NL_F64) return NL_ERR_ENUM;
  if (count < 0 || step < 1) return NL_ERR_SHAPE;
  if (count == 0) return NL_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !scratch) return NL_ERR_NULL;
  if (scratch_bytes < nl::clip_adam_scratch_bytes()) return NL_ERR_WORKSPACE;
  g_launches = max_norm > 0.0 ? 2 : 1;
  cudaError_t e = nl::launch_clip_adam(dtype, count, params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step,
                                       max_norm, write_clipped_grads, scratch, (cudaStream_t)stream);
  return e == cudaSuccess ? NL_OK : cuda_fail(e);
}

}  // extern "C"
