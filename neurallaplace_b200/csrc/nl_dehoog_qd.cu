// De Hoog quotient-difference recurrence between the tensor-core forward and backward kernels
// (reference: torchlaplace/inverse_laplace.py:737-866; the recurrence itself is nl_dehoog.cuh).
// One thread per (b, t, d) element, as BASELINE.json's north_star asks ("a per-thread recurrence that stays
// parallel over batch and time").  F and dL/dF are [d][k][t][b] complex planes, so the 32 threads of a warp
// (consecutive b) read and write 256 contiguous bytes per term.
#include <algorithm>
#include <cstdlib>

#include "nl_dehoog.cuh"
#include "nl_dehoog_reg.cuh"
#include "nl_kernels.h"

namespace nl {

namespace {
constexpr int kQdThreads = 128;

template <bool BWD>
__global__ void __launch_bounds__(kQdThreads) dehoog_qd_kernel(IltParams<float> P, int B, int T, int D,
                                                               const float* __restrict__ t,
                                                               const Cx<float>* __restrict__ F,
                                                               const float* __restrict__ gx, float* __restrict__ x,
                                                               Cx<float>* __restrict__ dF, Cx<float>* scratch,
                                                               int64_t stride) {
  const int n = P.n, M = (n - 1) / 2;
  const int64_t plane = (int64_t)T * B;
  const int64_t items = plane * D;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t item = slot; item < items; item += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(item / plane);
    const int64_t tb = item - (int64_t)d * plane;
    const int ti = (int)(tb / B);
    const int b = (int)(tb - (int64_t)ti * B);
    const TimeCtx<float> c = make_time_ctx<float>(P, t[ti], nullptr, 0);
    float sn, cs;
    m_sincos((float(kPi) * c.t) / c.T, &sn, &cs);  // z = exp(i*pi*t/T), inverse_laplace.py:842
    const Cx<float> z{cs, sn};
    const Cx<float>* Fp = F + (int64_t)d * n * plane + tb;
    auto a = [&](int k) { return Fp[(int64_t)k * plane]; };
    Scratch<float> scr{scratch, stride, slot};
    const int64_t xi = ((int64_t)b * T + ti) * D + d;
    if (!BWD) {
      const Cx<float> w = dehoog_forward<float>(M, z, a, scr);
      x[xi] = c.pref * w.re;
    } else {
      const float g = gx[xi] * c.pref;
      Cx<float>* dFp = dF + (int64_t)d * n * plane + tb;
      auto emit = [&](int k, Cx<float> dw) { dFp[(int64_t)k * plane] = Cx<float>{g * dw.re, -g * dw.im}; };
      dehoog_forward_backward<float>(M, z, a, emit, scr);
    }
  }
}

// ---- register-resident variant for 17 / 33 terms (nl_dehoog_reg.cuh) ----
struct TabRef {
  Cx<float>* base;
  int64_t stride, slot;
  __device__ __forceinline__ Cx<float>& operator()(int idx) const { return base[(int64_t)idx * stride + slot]; }
};
struct HistRef {
  Cx<float>* base;
  __device__ __forceinline__ Cx<float>& operator()(int idx) const { return base[idx * kQdThreads + threadIdx.x]; }
};

// per-thread ring in shared memory filled by cp.async (8 bytes per entry and thread; a thread only ever reads what
// it copied itself, so cp.async.wait_group is the only synchronisation)
struct StageRef {
  Cx<float>* ring;          // [kDhStageEntries][kQdThreads]
  const Cx<float>* tbase;   // this thread's table slot
  int64_t stride;
  __device__ __forceinline__ void issue(int k, int idx) const {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + k * kQdThreads + threadIdx.x);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(tbase + (int64_t)idx * stride) : "memory");
  }
  __device__ __forceinline__ void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  template <int N>
  __device__ __forceinline__ void wait() const {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
  }
  __device__ __forceinline__ Cx<float> get(int k) const { return ring[k * kQdThreads + threadIdx.x]; }
};

template <int M, bool BWD>
__global__ void __launch_bounds__(kQdThreads, BWD ? 2 : 3) dehoog_qd_reg_kernel(
    IltParams<float> P, int B, int T, int D, const float* __restrict__ t, const Cx<float>* __restrict__ F,
    const float* __restrict__ gx, float* __restrict__ x, Cx<float>* __restrict__ dF, Cx<float>* table,
    int64_t stride) {
  extern __shared__ __align__(16) unsigned char qd_smem[];
  constexpr int n = 2 * M + 1;
  const int64_t plane = (int64_t)T * B;
  const int64_t items = plane * D;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t item = slot; item < items; item += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(item / plane);
    const int64_t tb = item - (int64_t)d * plane;
    const int ti = (int)(tb / B);
    const int b = (int)(tb - (int64_t)ti * B);
    const TimeCtx<float> c = make_time_ctx<float>(P, t[ti], nullptr, 0);
    float sn, cs;
    m_sincos((float(kPi) * c.t) / c.T, &sn, &cs);
    const Cx<float> z{cs, sn};
    const Cx<float>* Fp = F + (int64_t)d * n * plane + tb;
    auto a = [&](int k) { return Fp[(int64_t)k * plane]; };
    const int64_t xi = ((int64_t)b * T + ti) * D + d;
    if (!BWD) {
      const Cx<float> w = dehoog_reg_forward<float, M>(z, a);
      x[xi] = c.pref * w.re;
    } else {
      const float g = gx[xi] * c.pref;
      Cx<float>* dFp = dF + (int64_t)d * n * plane + tb;
      auto emit = [&](int k, Cx<float> dw) { dFp[(int64_t)k * plane] = Cx<float>{g * dw.re, -g * dw.im}; };
      Cx<float>* sm = reinterpret_cast<Cx<float>*>(qd_smem);
      dehoog_reg_forward_backward<float, M>(
          z, a, emit, TabRef{table, stride, slot}, HistRef{sm},
          StageRef{sm + DehoogRegLayout<M>::kHistEntries * kQdThreads, table + slot, stride});
    }
  }
}

template <int M, bool BWD>
size_t reg_smem() {
  return BWD ? (size_t)(DehoogRegLayout<M>::kHistEntries + kDhStageEntries) * kQdThreads * sizeof(Cx<float>) : 0;
}
template <int M, bool BWD>
int reg_blocks() {
  static int per_sm = [] {
    int v = 0;
    cudaFuncSetAttribute(dehoog_qd_reg_kernel<M, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)reg_smem<M, BWD>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dehoog_qd_reg_kernel<M, BWD>, kQdThreads, reg_smem<M, BWD>());
    return v > 0 ? v : 1;
  }();
  return per_sm * device_sm_count();
}
bool reg_variant(const nl_desc* d) {
  static const bool off = [] { const char* v = getenv("NL_B200_DEHOOG_REG"); return v && atoi(v) == 0; }();
  return !off && (d->n == 17 || d->n == 33);
}
template <int M>
cudaError_t launch_reg(const nl_desc* d, bool bwd, const IltParams<float>& P, const void* t, const void* F,
                       const void* gx, void* x, void* dF, Cx<float>* scr, cudaStream_t st) {
  const int64_t items = (int64_t)d->B * d->T * d->D;
  const int64_t need = (items + kQdThreads - 1) / kQdThreads;
  if (bwd) {
    const int64_t blocks = std::min<int64_t>(need, reg_blocks<M, true>());
    dehoog_qd_reg_kernel<M, true><<<(unsigned)blocks, kQdThreads, reg_smem<M, true>(), st>>>(
        P, d->B, d->T, d->D, (const float*)t, (const Cx<float>*)F, (const float*)gx, nullptr, (Cx<float>*)dF, scr,
        blocks * kQdThreads);
  } else {
    const int64_t blocks = std::min<int64_t>(need, (int64_t)reg_blocks<M, false>() * 4);
    dehoog_qd_reg_kernel<M, false><<<(unsigned)blocks, kQdThreads, 0, st>>>(
        P, d->B, d->T, d->D, (const float*)t, (const Cx<float>*)F, nullptr, (float*)x, nullptr, nullptr, 0);
  }
  return cudaGetLastError();
}

int64_t qd_blocks(const nl_desc* d) {
  const int64_t items = (int64_t)d->B * d->T * d->D;
  const int64_t need = (items + kQdThreads - 1) / kQdThreads;
  return std::min<int64_t>(need, (int64_t)device_sm_count() * 8);
}
}  // namespace

size_t dehoog_qd_scratch_bytes(const nl_desc* d, bool bwd) {
  if (reg_variant(d)) {
    if (!bwd) return 256;
    const int blocks = d->n == 17 ? reg_blocks<8, true>() : reg_blocks<16, true>();
    const size_t ent = d->n == 17 ? DehoogRegLayout<8>::kTabEntries : DehoogRegLayout<16>::kTabEntries;
    return ent * (size_t)blocks * kQdThreads * sizeof(Cx<float>) + 256;
  }
  return (size_t)dehoog_scratch_entries(d->n, bwd) * (size_t)qd_blocks(d) * kQdThreads * sizeof(Cx<float>) + 256;
}

cudaError_t launch_dehoog_qd(const nl_desc* d, bool bwd, const void* t, const void* F, const void* gx, void* x,
                             void* dF, void* scratch, cudaStream_t st) {
  const int64_t blocks = qd_blocks(d);
  if (blocks == 0) return cudaSuccess;
  IltParams<float> P{};
  P.family = NL_DEHOOG;
  P.n = d->n;
  P.head = d->head;
  P.alpha = (float)d->alpha;
  P.log_tol = (float)d->log_tol;
  P.scale = (float)d->scale;
  Cx<float>* scr = (Cx<float>*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  if (reg_variant(d))
    return d->n == 17 ? launch_reg<8>(d, bwd, P, t, F, gx, x, dF, scr, st)
                      : launch_reg<16>(d, bwd, P, t, F, gx, x, dF, scr, st);
  const int64_t stride = blocks * kQdThreads;
  if (bwd)
    dehoog_qd_kernel<true><<<(unsigned)blocks, kQdThreads, 0, st>>>(P, d->B, d->T, d->D, (const float*)t,
                                                                   (const Cx<float>*)F, (const float*)gx, nullptr,
                                                                   (Cx<float>*)dF, scr, stride);
  else
    dehoog_qd_kernel<false><<<(unsigned)blocks, kQdThreads, 0, st>>>(P, d->B, d->T, d->D, (const float*)t,
                                                                    (const Cx<float>*)F, nullptr, (float*)x, nullptr,
                                                                    scr, stride);
  return cudaGetLastError();
}

}  // namespace nl
