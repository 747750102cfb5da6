// Training-step tail over ONE flat parameter buffer: global-norm gradient clip + Adam update in two launches
// (what the reference's training loops do with stock PyTorch after backward --
// torch.nn.utils.clip_grad_norm_(params, max_norm) then torch.optim.Adam.step(): examples/simple_demo.py:349-350,
// experiments/utils.py:70-71 -- ~10 small launches per parameter tensor there).  The flat buffer is the one the
// data-parallel all-reduce of the MLP gradients already uses (neurallaplace_b200/parallel.py), so a step is
// backward -> one all-reduce -> these two kernels.  HBM-bound elementwise work: 16 B read + 12 B written per
// parameter (+4 B when the clipped gradient is written back); deterministic (no atomics).
#include <algorithm>

#include "nl_kernels.h"

namespace nl {
namespace {
constexpr int kOptThreads = 256;
constexpr int kOptMaxBlocks = 1024;

template <typename R>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kOptThreads / 32; ++w) s += sh[w];  // same order in every thread
  __syncthreads();
  return s;
}

// partial[b] = sum of g^2 over the slice of block b (accumulated in double)
template <typename R>
__global__ void __launch_bounds__(kOptThreads) grad_sumsq_kernel(const R* __restrict__ g, int64_t count,
                                                                 double* __restrict__ partial) {
  __shared__ double sh[kOptThreads / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kOptThreads) {
    const double v = (double)g[i];
    acc += v * v;
  }
  const double s = block_sum<R>(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

template <typename R>
__global__ void __launch_bounds__(kOptThreads) clip_adam_kernel(R* __restrict__ p, R* __restrict__ g,
                                                                R* __restrict__ m, R* __restrict__ v, int64_t count,
                                                                const double* __restrict__ partial, int nblocks,
                                                                R lr, R beta1, R beta2, R eps, R bc1, R bc2_sqrt,
                                                                double max_norm, int write_grad,
                                                                double* __restrict__ norm_out) {
  __shared__ double sh[kOptThreads / 32];
  R coef = R(1);
  if (max_norm > 0.0) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kOptThreads) acc += partial[i];
    const double total = sqrt(block_sum<R>(acc, sh));
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const double c = max_norm / (total + 1e-6);
    coef = (R)(c < 1.0 ? c : 1.0);
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = total;
  }
  const R step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kOptThreads) {
    const R gi = g[i] * coef;
    // torch.optim.Adam (no amsgrad, no weight decay): exp_avg.lerp_(grad, 1 - beta1);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2); denom = sqrt(v)/sqrt(bc2) + eps
    const R mi = m[i] + (gi - m[i]) * (R(1) - beta1);
    const R vi = v[i] * beta2 + (R(1) - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const R denom = sqrt(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    if (write_grad) g[i] = gi;
  }
}

template <typename R>
cudaError_t launch_t(int64_t count, void* p, void* g, void* m, void* v, double lr, double beta1, double beta2,
                     double eps, int64_t step, double max_norm, int write_grad, void* scratch, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  double* partial = (double*)scratch;
  double* norm_out = partial + kOptMaxBlocks;
  const int blocks = (int)std::min<int64_t>((count + kOptThreads - 1) / kOptThreads,
                                            std::min(kOptMaxBlocks, device_sm_count() * 4));
  if (max_norm > 0.0) {
    grad_sumsq_kernel<R><<<blocks, kOptThreads, 0, st>>>((const R*)g, count, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  clip_adam_kernel<R><<<blocks, kOptThreads, 0, st>>>((R*)p, (R*)g, (R*)m, (R*)v, count, partial, blocks, (R)lr,
                                                     (R)beta1, (R)beta2, (R)eps, (R)bc1, (R)sqrt(bc2), max_norm,
                                                     write_grad, norm_out);
  return cudaGetLastError();
}
}  // namespace

size_t clip_adam_scratch_bytes() { return (kOptMaxBlocks + 1) * sizeof(double); }

cudaError_t launch_clip_adam(int dtype, int64_t count, void* p, void* g, void* m, void* v, double lr, double beta1,
                             double beta2, double eps, int64_t step, double max_norm, int write_grad, void* scratch,
                             cudaStream_t st) {
  return dtype == NL_F64 ? launch_t<double>(count, p, g, m, v, lr, beta1, beta2, eps, step, max_norm, write_grad,
                                            scratch, st)
                         : launch_t<float>(count, p, g, m, v, lr, beta1, beta2, eps, step, max_norm, write_grad,
                                           scratch, st);
}

}  // namespace nl
