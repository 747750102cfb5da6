// Internal launcher declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/nl_b200.h"

namespace nl {

template <typename R>
cudaError_t launch_query_points(const nl_desc* d, int64_t rows, const void* t, const void* Tper, const void* beta,
                                void* s_out, void* feat_out, cudaStream_t st);
template <typename R>
cudaError_t launch_sphere(int which, int64_t count, const void* a, const void* b, void* c, void* e, cudaStream_t st);
template <typename R>
cudaError_t launch_line_integrate(const nl_desc* d, int layout, bool bwd, const void* fa, const void* fb,
                                  const void* t, const void* Tper, const void* eta, const void* gx, void* x,
                                  void* ga, void* gb, cudaStream_t st);

void set_thread_scratch(void* ptr, size_t bytes);
size_t thread_scratch_bytes();
size_t line_integrate_scratch_bytes(const nl_desc* d, bool bwd);

// ---- fused SIMT path (any dtype / family / t_mode; nl_fused_simt.cu) ----
struct SimtPlan {
  int ok;            // 0 => shape not supported by the fused SIMT kernels
  int rp;            // rows per lane (tile rows = 32*rp)
  int BB, TT;        // tile geometry
  int grid;          // persistent CTAs
  int threads;       // threads per CTA
  size_t smem;       // dynamic shared memory bytes
  size_t slab_len;   // reals per CTA gradient slab (backward)
  size_t ws_bytes;   // workspace bytes
};
SimtPlan simt_plan(const nl_desc* d, int pass);
template <typename R>
cudaError_t launch_fused_simt(const nl_desc* d, bool bwd, const void* p, const void* t, const void* const* weights,
                              const void* beta, const void* eta, void* x, const void* gx, void* const* grads,
                              void* ws, cudaStream_t st, int* launches);

// ---- fused tcgen05 path (f32, canonical H=64; nl_fused_tc.cu) ----
struct TcPlan {
  int ok;
  int grid;
  size_t smem;
  size_t ws_bytes;
  int co_resident;  // forward: two CTAs fit on one SM (shared memory <= half)
  int ring;         // W3 is streamed through a ring (does not fit resident)
};
TcPlan tc_plan(const nl_desc* d, int pass);
// `saved`: optional activation stash of tc_saved_bytes(d) bytes -- written by the forward pass, consumed by the
// backward pass (which then skips the recomputation of layers 1-2); NULL = plain forward / recomputing backward.
cudaError_t launch_fused_tc(const nl_desc* d, bool bwd, const void* p, const void* t, const void* const* weights,
                            const void* beta, const void* eta, void* x, const void* gx, void* const* grads, void* ws,
                            cudaStream_t st, int* launches, void* saved = nullptr);
size_t tc_saved_bytes(const nl_desc* d);
// per-row time grids on the tensor-core path (point_layer1_kernel + the forward with h1 streamed from the stash)
TcPlan tc_pr_plan(const nl_desc* d, int pass);
size_t tc_pr_saved_bytes(const nl_desc* d);
cudaError_t launch_fused_tc_per_row(const nl_desc* d, bool bwd, const void* p, const void* t,
                                    const void* const* weights, const void* beta, const void* eta, void* x,
                                    const void* gx, void* const* grads, void* ws, cudaStream_t st, int* launches,
                                    void* saved);

// two-stream warp-specialised variant (nl_fused_tc2.cu); preferred when its smem / TMEM plan fits
TcPlan tc2_plan(const nl_desc* d, int pass);
cudaError_t launch_fused_tc2(const nl_desc* d, bool bwd, const void* p, const void* t, const void* const* weights,
                             const void* beta, const void* eta, void* x, const void* gx, void* const* grads, void* ws,
                             cudaStream_t st, int* launches);

// ---- De Hoog on the tensor-core path: F / dF cross HBM as [d][k][t][b] complex planes, the quotient-difference
// recurrence runs one thread per (b, t, d) in between (nl_dehoog_qd.cu) ----
TcPlan tc_dh_plan(const nl_desc* d, int pass);
size_t tc_dh_saved_bytes(const nl_desc* d);
cudaError_t launch_fused_tc_dehoog(const nl_desc* d, bool bwd, const void* p, const void* t,
                                   const void* const* weights, void* x, const void* gx, void* const* grads, void* ws,
                                   cudaStream_t st, int* launches, void* saved);
size_t dehoog_qd_scratch_bytes(const nl_desc* d, bool bwd);
cudaError_t launch_dehoog_qd(const nl_desc* d, bool bwd, const void* t, const void* F, const void* gx, void* x,
                             void* dF, void* scratch, cudaStream_t st, const void* Tper = nullptr);

// ---- fixed-contour De Hoog: one quotient-difference table per contour row, Pade recurrence per time point
// (nl_dehoog_fixed.cu) ----
size_t dehoog_fixed_scratch_bytes(int dtype, int64_t rows, int n, int64_t T, bool bwd);
cudaError_t launch_dehoog_fixed(int dtype, int64_t rows, int n, int64_t T, bool bwd, const void* F, const void* t,
                                double contour_T, double gamma, const void* gx, void* x, void* dF, void* scratch,
                                cudaStream_t st);

// ---- training-step tail over one flat buffer: global-norm clip + Adam (nl_optim.cu) ----
size_t clip_adam_scratch_bytes();
cudaError_t launch_clip_adam(int dtype, int64_t count, void* p, void* g, void* m, void* v, double lr, double beta1,
                             double beta2, double eps, int64_t step, double max_norm, int write_grad, void* scratch,
                             cudaStream_t st);

int device_sm_count();

}  // namespace nl
