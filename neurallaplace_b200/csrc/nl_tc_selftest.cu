// Debug entry point: one bf16x3 UMMA GEMM through the exact operand images / descriptors the fused
// tensor-core kernels use, so descriptor conventions can be validated in isolation on a GPU
// (tests/test_gpu_tc.py).  Not part of the public C ABI (include/nl_b200.h).
#include <cstdio>

#include "nl_tc_common.cuh"

namespace nl {
namespace tc {

// A: fp32 [rowsA][colsA] row-major, B: fp32 [rowsB][colsB] row-major (cols % 8 == 0, rows <= 128).
// a_mn = 0: A is K-major  (M = rowsA = 128, K = colsA)        D[m][n] = sum_k A[m][k]  * Bop[n][k]
// a_mn = 1: A is MN-major (M given,         K = rowsA)        D[m][n] = sum_k A[k][m]  * Bop[n][k]
// b_mn = 0: B is K-major  (N = rowsB, K = colsB)   Bop[n][k] = B[n][k]
// b_mn = 1: B is MN-major (N = given, K = rowsB)   Bop[n][k] = B[k][n]
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, int rowsA, int colsA, int a_mn, const float* __restrict__ B,
                     int rowsB, int colsB, int b_mn, int M, int N, int Kdim, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t a_bytes = image_bytes(rowsA, colsA), b_bytes = image_bytes(rowsB, colsB);
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  // zero everything first (including the slack the M=128 MN-major read runs into)
  for (int i = tid; i < (int)((2 * a_bytes + 2 * b_bytes + 65536) / 16); i += 128)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < rowsA)
    for (int c0 = 0; c0 < colsA; c0 += 8) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * colsA + c0 + i];
      store_chunk8(a_hi, a_lo, rowsA, tid, c0, v);
    }
  if (tid < rowsB)
    for (int c0 = 0; c0 < colsB; c0 += 8) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = B[(size_t)tid * colsB + c0 + i];
      store_chunk8(b_hi, b_lo, rowsB, tid, c0, v);
    }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    GemmOp g;
    if (a_mn) operand_mnmajor(smem_u32(a_hi), (uint32_t)a_bytes, rowsA, &g.a_hi, &g.a_lo_off, &g.a_lbo, &g.a_sbo, &g.a_step);
    else      operand_kmajor(smem_u32(a_hi), (uint32_t)a_bytes, rowsA, &g.a_hi, &g.a_lo_off, &g.a_lbo, &g.a_sbo, &g.a_step);
    if (b_mn) operand_mnmajor(smem_u32(b_hi), (uint32_t)b_bytes, rowsB, &g.b_hi, &g.b_lo_off, &g.b_lbo, &g.b_sbo, &g.b_step);
    else      operand_kmajor(smem_u32(b_hi), (uint32_t)b_bytes, rowsB, &g.b_hi, &g.b_lo_off, &g.b_lbo, &g.b_sbo, &g.b_step);
    g.idesc = make_idesc_bf16(M, N, a_mn, b_mn);
    g.ksteps = Kdim / 16;
    issue_gemm(g, tmem, false);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // read back: thread <-> TMEM lane
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  int row;
  bool active;
  if (M == 128) {
    row = tid;
    active = true;
  } else {  // M == 64: rows 16w..16w+15 live in lanes 32w..32w+15
    row = warp * 16 + lane;
    active = lane < 16;
  }
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tmem_ld8(lane_addr + c0, v);
    tmem_ld_wait();
    if (active)
      for (int i = 0; i < 8; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace nl

extern "C" int nl_debug_umma_selftest(const float* A, int rowsA, int colsA, int a_mn, const float* B, int rowsB,
                                      int colsB, int b_mn, int M, int N, int Kdim, float* D, void* stream) {
  using namespace nl::tc;
  if (rowsA > 128 || rowsB > 128 || colsA % 8 || colsB % 8 || Kdim % 16 || (M != 64 && M != 128) || N % 8 || N > 256)
    return -2;
  size_t smem = 2 * image_bytes(rowsA, colsA) + 2 * image_bytes(rowsB, colsB) + 65536 + 1024;
  if (smem > 227 * 1024) return -2;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -5;
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, rowsA, colsA, a_mn, B, rowsB, colsB, b_mn, M, N,
                                                              Kdim, D);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
