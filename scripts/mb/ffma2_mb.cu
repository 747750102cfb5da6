// Microbenchmark: issue cost of packed FP32 (FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2) against scalar FFMA on sm_100a,
// alone and next to MUFU work -- the fused kernels' epilogues are issue-slot / MUFU bound, not FMA-pipe bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_mb ffma2_mb.cu && ./ffma2_mb
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mode 0: 8 independent scalar FFMA chains; 1: 8 independent FFMA2 chains; 2: 8 FFMA + 2 MUFU; 3: 8 FFMA2 + 2 MUFU;
// 4: 2 MUFU only
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float a[8];
  unsigned long long p[8];
  float m0 = threadIdx.x * 1e-3f, m1 = m0 + 0.5f;
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 0.001f + i;
    p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f);
  }
  const float b = 0.999f, c = 0.001f;
  const unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  const unsigned long long pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], b, c);
    }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pb, pc);
    }
    if (MODE >= 2) {
      m0 = ex2a(m0);
      m1 = ex2a(m1);
    }
  }
  const long long t1 = clock64();
  float s = m0 + m1;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 8);
  const int iters = 20000;
  k<MODE><<<148, warps_per_sm * 32>>>(out, iters, clk);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps_per_sm * 32>>>(out, iters, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per_iter = (double)h / iters;
  printf("%-28s warps/SM %2d: %7.2f clk per iteration per SM-wide pass (%.2f clk per warp-iteration per SMSP)\n", name,
         warps_per_sm, per_iter, per_iter / (warps_per_sm / 4.0));
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("8 x FFMA", w);
    run<1>("8 x FFMA2", w);
    run<4>("2 x MUFU.EX2", w);
    run<2>("8 x FFMA  + 2 x MUFU.EX2", w);
    run<3>("8 x FFMA2 + 2 x MUFU.EX2", w);
  }
  return 0;
}
