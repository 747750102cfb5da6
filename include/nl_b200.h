/*
 * nl_b200.h -- C ABI of the B200-native torchlaplace reconstruction hot path.
 *
 * The reference (samholt/NeuralLaplace, package `torchlaplace`) is pure Python/PyTorch and has
 * no FFI layer of its own; its boundary for this path is the Python API
 *     torchlaplace.laplace_reconstruct                      (torchlaplace/core.py:25-150)
 *     torchlaplace.inverse_laplace.{Fourier,DeHoog,CME,FixedTablot,Stehfest,Euler,Gaver}
 *         .compute_s / .line_integrate* / .forward          (torchlaplace/inverse_laplace.py:225-1689)
 *     torchlaplace.transformations.*                        (torchlaplace/transformations.py:28-151)
 * Each entry point below names the reference function(s) it replaces.  The Python shim that
 * binds them (ctypes) lives in neurallaplace_b200/_lib.py; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous row-major array unless named *_host;
 *   - real arrays are float (dtype NL_F32) or double (NL_F64); complex arrays are interleaved
 *     (re, im) pairs of that real type (the storage of torch.cfloat / torch.cdouble);
 *   - weights use the nn.Linear layout W[out][in];
 *   - functions return 0 or a negative nl_status, never throw, never allocate or free caller
 *     memory, never synchronise the stream; they may be called from any host thread;
 *   - the caller passes the CUDA stream (cudaStream_t as void*) and a workspace of at least
 *     nl_workspace_bytes() bytes; the library writes the workspace (no pre-zeroing needed).
 */
#ifndef NL_B200_H_
#define NL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NL_B200_ABI_VERSION 3

typedef enum nl_status {
  NL_OK = 0,
  NL_ERR_NULL = -1,        /* a required pointer was NULL                                   */
  NL_ERR_SHAPE = -2,       /* a dimension is zero/negative or out of the supported range     */
  NL_ERR_UNSUPPORTED = -3, /* valid request the fused kernels do not cover (use split path)   */
  NL_ERR_WORKSPACE = -4,   /* workspace smaller than nl_workspace_bytes()                    */
  NL_ERR_CUDA = -5,        /* a CUDA runtime call failed (see nl_last_cuda_error())          */
  NL_ERR_ENUM = -6         /* invalid family / dtype / mode value                            */
} nl_status;

enum { NL_F32 = 0, NL_F64 = 1 };
/* ILT families as the kernels see them.
 *   NL_FOURIER : s_k = gamma(t) + i*pi*k/T, weighted sum with twiddles (Fourier, l.1045-1306)
 *   NL_AW      : Abate-Whitt form s_k = beta_k/t, x = Re(sum eta_k F_k)/t
 *                (CME l.1581-1689, Euler l.1420-1578, Gaver l.1309-1417, FixedTablot l.225-408,
 *                 Stehfest l.411-511 -- only the host-side beta/eta tables differ)
 *   NL_DEHOOG  : Fourier nodes + quotient-difference / Pade acceleration (DeHoog, l.514-1042)  */
enum { NL_FOURIER = 0, NL_AW = 1, NL_DEHOOG = 2 };
enum { NL_T_SHARED = 0, NL_T_PER_ROW = 1 };   /* t[T] shared by the batch, or t[B][T]        */
enum { NL_HEAD_SPHERE = 0, NL_HEAD_RAW = 1 }; /* use_sphere_projection True / False           */
/* layout of F handed to the stand-alone line-integrate kernels */
enum { NL_F_SPHERE_PLANAR = 0, NL_F_COMPLEX_PLANAR = 1, NL_F_COMPLEX_INTERLEAVED = 2 };

typedef struct nl_desc {
  int32_t B;      /* batch rows (p.shape[0])                                               */
  int32_t T;      /* time points per row                                                   */
  int32_t K;      /* latent dim (p.shape[1])                                               */
  int32_t H;      /* hidden units of the canonical laplace_rep_func MLP                    */
  int32_t D;      /* recon_dim / d_obs                                                     */
  int32_t n;      /* number of s points actually used per time point                       */
  int32_t family; /* NL_FOURIER / NL_AW / NL_DEHOOG                                        */
  int32_t t_mode; /* NL_T_SHARED / NL_T_PER_ROW                                            */
  int32_t head;   /* NL_HEAD_SPHERE / NL_HEAD_RAW                                          */
  int32_t dtype;  /* NL_F32 / NL_F64                                                       */
  int32_t impl;   /* 0 = auto, 1 = force SIMT kernels, 2 = force tcgen05 kernels           */
  int32_t reserved;
  /* Fourier / DeHoog constants, ALREADY float32-rounded where the reference rounds them
   * (alpha, log(tol), scale are float32 tensors in the reference even on the double path,
   * inverse_laplace.py:1107-1115, 555-563).                                                */
  double alpha;
  double log_tol;
  double scale;
} nl_desc;

/* ---- bookkeeping ------------------------------------------------------------------- */
int nl_abi_version(void);
const char* nl_strerror(int status);
const char* nl_last_cuda_error(void);
/* Workspace needed by nl_reconstruct_forward (pass=0) / nl_reconstruct_backward (pass=1). */
size_t nl_workspace_bytes(const nl_desc* desc, int pass);
/* Which kernel family `impl=auto` would pick: 1 = SIMT, 2 = tcgen05.                       */
int nl_selected_impl(const nl_desc* desc, int pass);
/* Number of kernel launches the last nl_reconstruct_* call on this thread issued.          */
int nl_last_launch_count(void);

/* ---- fused hot path -----------------------------------------------------------------
 * Replaces core.py:104-150 with the canonical laplace_rep_func
 * (tests/test_laplace_rep_func.py:36-64; Sequential(Linear,Tanh,Linear,Tanh,Linear) + tanh
 * sphere heads): s-point generation (compute_s), complex_to_spherical_riemann, the MLP,
 * spherical_to_complex and line_integrate_all_multi[_batch_time] in ONE kernel.
 *   p[B][K], t[T] or t[B][T], weights = {W1[H][2n+K], b1[H], W2[H][H], b2[H], W3[2Dn][H], b3[2Dn]}
 *   beta[n], eta[n] complex (NL_AW only, else may be NULL);  x[B][T][D].                     */
int nl_reconstruct_forward(const nl_desc* desc, const void* p, const void* t,
                           const void* const* weights, const void* beta, const void* eta, void* x,
                           void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Backward of the above (replaces autograd over the same op chain): recomputes the forward
 * per tile and returns d/d{W1,b1,W2,b2,W3,b3,p} of sum(grad_x * x).
 *   grads = {dW1, db1, dW2, db2, dW3, db3, dp}; every array is OVERWRITTEN.                  */
int nl_reconstruct_backward(const nl_desc* desc, const void* p, const void* t,
                            const void* const* weights, const void* beta, const void* eta,
                            const void* grad_x, void* const* grads, void* workspace,
                            size_t workspace_bytes, void* cuda_stream);

/* Training variant of the pair above (what autograd's saved-tensor mechanism does for the reference's
 * op chain -- every intermediate of core.py:104-150 is kept for backward; here only the two hidden
 * activations are).  nl_reconstruct_forward_save additionally writes the hidden activations h1, h2 of
 * the laplace_rep_func MLP (as the bf16 hi/lo tensor-core operand images, 512 B per (b,t) point) into
 * `saved` (nl_saved_bytes(desc) bytes, caller-owned, must stay untouched until the backward call);
 * nl_reconstruct_backward_saved streams them back instead of recomputing layers 1-2.  Results are
 * bit-identical to nl_reconstruct_forward / within rounding of nl_reconstruct_backward.
 * nl_saved_bytes returns 0 when the shape is not covered (use the plain pair).                     */
size_t nl_saved_bytes(const nl_desc* desc);
int nl_reconstruct_forward_save(const nl_desc* desc, const void* p, const void* t,
                                const void* const* weights, const void* beta, const void* eta,
                                void* x, void* saved, size_t saved_bytes, void* workspace,
                                size_t workspace_bytes, void* cuda_stream);
int nl_reconstruct_backward_saved(const nl_desc* desc, const void* p, const void* t,
                                  const void* const* weights, const void* beta, const void* eta,
                                  const void* grad_x, const void* saved, size_t saved_bytes,
                                  void* const* grads, void* workspace, size_t workspace_bytes,
                                  void* cuda_stream);

/* ---- split path (arbitrary user laplace_rep_func, and the class-level ILT API) -------
 * nl_query_points: <ILT>.compute_s (inverse_laplace.py:316-344, 494-496, 703-735, 1122-1147,
 * 1640-1643) for `rows` time points, optionally followed by complex_to_spherical_riemann
 * (transformations.py:56-89).  Tper may be NULL (T = scale * t); s_out[rows][n] complex and
 * feat_out[rows][2n] (= [theta_s | phi_s] or [Re s | Im s] per desc->head) may each be NULL.
 * Only desc->{n,family,head,dtype,alpha,log_tol,scale} are read.                            */
int nl_query_points(const nl_desc* desc, int64_t rows, const void* t, const void* Tper,
                    const void* beta, void* s_out, void* feat_out, void* cuda_stream);

/* nl_line_integrate_{forward,backward}: <ILT>.line_integrate / line_integrate_multi /
 * line_integrate_all_multi / line_integrate_all_multi_batch_time (inverse_laplace.py:346-408,
 * 498-502, 737-895, 1149-1267, 1645-1683), optionally fused with spherical_to_complex
 * (transformations.py:95-121) when f_layout == NL_F_SPHERE_PLANAR.
 *   fa, fb: [B][T][D][n] planar pair (theta,phi) or (re,im); for NL_F_COMPLEX_INTERLEAVED fa is
 *   the complex array and fb is ignored.  t and Tper follow desc->t_mode ([T] or [B][T]);
 *   Tper may be NULL (T = scale * t for Fourier/DeHoog, T = t for AW).  x[B][T][D].
 *   backward writes d/dfa, d/dfb (same layout as the inputs) of sum(grad_x * x).             */
int nl_line_integrate_forward(const nl_desc* desc, int f_layout, const void* fa, const void* fb,
                              const void* t, const void* Tper, const void* eta, void* x,
                              void* cuda_stream);
int nl_line_integrate_backward(const nl_desc* desc, int f_layout, const void* fa, const void* fb,
                               const void* t, const void* Tper, const void* eta,
                               const void* grad_x, void* grad_fa, void* grad_fb,
                               void* cuda_stream);

/* Scratch for nl_line_integrate_* (they have no workspace argument).  Only the De Hoog family needs any
 * (F / dL/dF planes + the recurrence's table): nl_line_integrate_scratch_bytes says how much (0 otherwise).
 * nl_set_scratch registers a caller-owned device buffer for the calling host thread (NULL clears it); a
 * De Hoog line-integrate call without a large enough registered buffer returns NL_ERR_WORKSPACE (the library
 * never allocates).                                                                                        */
size_t nl_line_integrate_scratch_bytes(const nl_desc* desc, int pass);
void nl_set_scratch(void* device_ptr, size_t bytes);

/* ---- fixed-contour De Hoog ----------------------------------
 * Replaces DeHoog.fixed_line_integrate (inverse_laplace.py:587-701): ONE contour F[rows][n] complex (the
 * Laplace representation at the s points of DeHoog.compute_fixed_s(time_max), l.567-585), evaluated at every
 * t[T]:  x[row][i] = exp(gamma t_i) / contour_T * Re(A / B)(z_i),  z_i = exp(i pi t_i / contour_T).
 * The quotient-difference table is built once per row, only the Pade recurrence runs per time point.
 *   contour_T = scale * time_max and gamma = alpha - log(tol) / (scale * contour_T), both ALREADY rounded
 *   to float32 as the reference computes them there (l.640-642), passed as doubles.
 *   backward writes grad_F[rows][n] (complex interleaved: d/dRe, d/dIm) of sum(grad_x * x), grad_x[rows][T].
 *   scratch: nl_dehoog_fixed_scratch_bytes(dtype, rows, n, T, pass) bytes, caller-owned (pass 0 fwd, 1 bwd). */
size_t nl_dehoog_fixed_scratch_bytes(int dtype, int64_t rows, int n, int64_t T, int pass);
int nl_dehoog_fixed_forward(int dtype, int64_t rows, int n, int64_t T, const void* F, const void* t,
                            double contour_T, double gamma, void* x, void* scratch, size_t scratch_bytes,
                            void* cuda_stream);
int nl_dehoog_fixed_backward(int dtype, int64_t rows, int n, int64_t T, const void* F, const void* t,
                             double contour_T, double gamma, const void* grad_x, void* grad_F, void* scratch,
                             size_t scratch_bytes, void* cuda_stream);

/* ---- sphere maps (transformations.py:28-151), elementwise over `count` values -------- */
int nl_complex_to_sphere(int dtype, int64_t count, const void* s_re, const void* s_im, void* theta,
                         void* phi, void* cuda_stream);
int nl_sphere_to_complex(int dtype, int64_t count, const void* theta, const void* phi, void* s_re,
                         void* s_im, void* cuda_stream);

/* ---- training-step tail ------------------------------------
 * Global-norm gradient clip + Adam update over ONE flat buffer of `count` parameters, two launches:
 * replaces torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step()
 * (examples/simple_demo.py:349-350, experiments/utils.py:70-71; Adam without amsgrad / weight decay).
 *   params, grads, exp_avg, exp_avg_sq: `count` values each; params / exp_avg / exp_avg_sq are updated in
 *   place; grads are read, and overwritten with the clipped gradients when write_clipped_grads != 0
 *   (clip_grad_norm_ scales them in place).  step = 1, 2, ... (bias correction).  max_norm <= 0: no clip.
 *   scratch: nl_clip_adam_scratch_bytes() bytes; the gradient norm before clipping is left in the double at
 *   byte offset 8192 of it.                                                                          */
size_t nl_clip_adam_scratch_bytes(void);
int nl_clip_adam_step(int dtype, int64_t count, void* params, void* grads, void* exp_avg,
                      void* exp_avg_sq, double lr, double beta1, double beta2, double eps, int64_t step,
                      double max_norm, int write_clipped_grads, void* scratch, size_t scratch_bytes,
                      void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NL_B200_H_ */
