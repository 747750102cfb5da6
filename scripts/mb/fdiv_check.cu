// Checks the claim in nl_dehoog_sm.cuh: the branch-free quotient chain (rcp.approx + Newton + residual) returns
// the same float as the IEEE division whenever both operands have exponents within [-60, 60].
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../neurallaplace_b200/csrc fdiv_check.cu -o fdiv_check
#include <cstdio>
#include <cstdint>
#include "nl_dehoog_sm.cuh"

__device__ unsigned long long g_bad, g_maxulp, g_cbad;

__device__ uint32_t rng(uint64_t& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (uint32_t)(s >> 33) ^ (uint32_t)s;
}
__device__ float rnd_float(uint64_t& s) {
  const uint32_t m = rng(s) & 0x7fffffu, e = 67u + rng(s) % 121u, sg = rng(s) & 1u;  // exponent field 67..187
  return __uint_as_float((sg << 31) | (e << 23) | m);
}
__global__ void check(int iters) {
  uint64_t s = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long bad = 0, maxulp = 0, cbad = 0;
  for (int i = 0; i < iters; ++i) {
    const float a = rnd_float(s), b = rnd_float(s);
#ifdef __CUDA_ARCH__
    const float q0 = a / b, q1 = nl::nl_fdiv_chain(a, b);
#else
    const float q0 = a / b, q1 = q0;
#endif
    if (__float_as_uint(q0) != __float_as_uint(q1)) {
      ++bad;
      const long long d = (long long)__float_as_int(q0) - (long long)__float_as_int(q1);
      const unsigned long long u = d < 0 ? -d : d;
      if (u > maxulp) maxulp = u;
    }
    // the batched complex division against Smith's algorithm with IEEE divisions
    nl::Cx<float> x[4], y[4], o[4];
    for (int k = 0; k < 4; ++k) {
      x[k] = {rnd_float(s), rnd_float(s)};
      y[k] = {rnd_float(s), rnd_float(s)};
    }
    nl::cdiv_batch<float, 4>(x, y, o);
    for (int k = 0; k < 4; ++k) {
      const nl::Cx<float> r = nl::cdiv(x[k], y[k]);
      if (__float_as_uint(r.re) != __float_as_uint(o[k].re) || __float_as_uint(r.im) != __float_as_uint(o[k].im)) ++cbad;
    }
  }
  atomicAdd(&g_bad, bad);
  atomicMax(&g_maxulp, maxulp);
  atomicAdd(&g_cbad, cbad);
}
int main() {
  unsigned long long z = 0;
  cudaMemcpyToSymbol(g_bad, &z, 8); cudaMemcpyToSymbol(g_maxulp, &z, 8); cudaMemcpyToSymbol(g_cbad, &z, 8);
  const int blocks = 592, threads = 256, iters = 2000;
  check<<<blocks, threads>>>(iters);
  cudaDeviceSynchronize();
  unsigned long long bad, mu, cb;
  cudaMemcpyFromSymbol(&bad, g_bad, 8); cudaMemcpyFromSymbol(&mu, g_maxulp, 8); cudaMemcpyFromSymbol(&cb, g_cbad, 8);
  const double n = (double)blocks * threads * iters;
  printf("real quotients: %.3g pairs, %llu differ from a/b (max %llu ulp)\n", n, bad, mu);
  printf("complex quotients (cdiv_batch vs cdiv): %.3g, %llu differ\n", 4 * n, cb);
  return 0;
}
