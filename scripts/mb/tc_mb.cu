// Micro-benchmark (development aid, not part of the product): cost of tcgen05.mma dispatches in the
// exact operand layouts of the fused kernels.  One elected thread issues `reps` bf16x3 GEMMs back to
// back and the CTA measures clocks from first issue to completion (commit -> mbarrier).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../neurallaplace_b200/csrc/nl_tc_common.cuh"
using namespace nl::tc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Cfg { int a_rows, a_mn, b_rows, b_mn, M, N, K, a_tmem, reps, stage, mode; };
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) mb_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00, 0x3c003c00, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t_issue = 0, t_done = 0;
  uint32_t ph = 0;
  bool issuer = false;
  if (MODE == 0) issuer = (tid == 0);
  else if (warp == 0) issuer = elect_one() != 0;
  if (issuer) {
    GemmOp g;
    const uint32_t a_img = smem_u32(smem), b_img = smem_u32(smem + 96 * 1024);
    if (c.a_mn) operand_mnmajor(a_img, 40960, c.a_rows, &g.a_hi, &g.a_lo_off, &g.a_lbo, &g.a_sbo, &g.a_step);
    else        operand_kmajor(a_img, 40960, c.a_rows, &g.a_hi, &g.a_lo_off, &g.a_lbo, &g.a_sbo, &g.a_step);
    if (c.b_mn) operand_mnmajor(b_img, 40960, c.b_rows, &g.b_hi, &g.b_lo_off, &g.b_lbo, &g.b_sbo, &g.b_step);
    else        operand_kmajor(b_img, 40960, c.b_rows, &g.b_hi, &g.b_lo_off, &g.b_lbo, &g.b_sbo, &g.b_step);
    g.idesc = make_idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
    g.ksteps = c.K / 16;
    const long long t0 = clock64();
    if (c.stage) {
      // stage mode: each GEMM is issued, committed and waited for (dependent chain latency)
      for (int r = 0; r < c.reps; ++r) {
        issue_gemm(g, tmem + 256, false);
        umma_commit(&bar);
        mbar_wait(&bar, ph); ph ^= 1;
        tc_fence_after();
      }
      t_issue = t_done = clock64() - t0;
    } else {
      for (int r = 0; r < c.reps; ++r) {
        if (c.a_tmem) {
          const uint64_t b_inc = (uint64_t)(g.b_step >> 4);
          uint32_t acc = 0;
          for (int pass = 0; pass < 3; ++pass) {
            uint64_t bd = make_sdesc(g.b_hi + (pass == 1 ? g.b_lo_off : 0), g.b_lbo, g.b_sbo);
            for (int ks = 0; ks < g.ksteps; ++ks) {
              umma_bf16_ts(tmem + 256, tmem + (pass == 2 ? 64 : 0) + ks * 8, bd, g.idesc, acc);
              acc = 1; bd += b_inc;
            }
          }
        } else {
          issue_gemm(g, tmem + 256, false);
        }
      }
      t_issue = clock64() - t0;
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t_done = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (issuer && blockIdx.x == 0) { out[0] = t_issue; out[1] = t_done; }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mb_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  const int R = 200;
  std::vector<Named> cases = {
      {"L2   A k128 B k64  M128 N64 K64 ", {128, 0, 64, 0, 128, 64, 64, 0, R, 0, 0}},
      {"L3   A k128 B k80  M128 N80 K64 ", {128, 0, 80, 0, 128, 80, 64, 0, R, 0, 0}},
      {"L3w  A k128 B k128 M128 N128 K64", {128, 0, 128, 0, 128, 128, 64, 0, R, 0, 0}},
      {"L3x  A k128 B k256 M128 N256 K64", {128, 0, 128, 0, 128, 256, 64, 0, R, 0, 0}},
      {"dH2  A k128 B mn80 M128 N64 K80 ", {128, 0, 80, 1, 128, 64, 80, 0, R, 0, 0}},
      {"dW3  A mn128 B mn128 M128 N80 K128", {128, 1, 128, 1, 128, 80, 128, 0, R, 0, 0}},
      {"dW2  A mn128 B mn128 M64 N72 K128", {128, 1, 128, 1, 64, 72, 128, 0, R, 0, 0}},
      {"dW2b A mn128 B mn128 M128 N72 K128", {128, 1, 128, 1, 128, 72, 128, 0, R, 0, 0}},
      {"D1   A mn128 B mn128 M64 N8 K128", {128, 1, 128, 1, 64, 8, 128, 0, R, 0, 0}},
      {"D1b  A mn128 B mn128 M128 N8 K128", {128, 1, 128, 1, 128, 8, 128, 0, R, 0, 0}},
      {"L2ts A tmem  B k64  M128 N64 K64 ", {128, 0, 64, 0, 128, 64, 64, 1, R, 0, 0}},
      {"L3ts A tmem  B k80  M128 N80 K64 ", {128, 0, 80, 0, 128, 80, 64, 1, R, 0, 0}},
      {"L2 stage (issue+commit+wait)", {128, 0, 64, 0, 128, 64, 64, 0, R, 1, 0}},
      {"L3 stage", {128, 0, 80, 0, 128, 80, 64, 0, R, 1, 0}},
      {"D1 stage", {128, 1, 128, 1, 64, 8, 128, 0, R, 1, 0}},
      {"dW3 stage", {128, 1, 128, 1, 128, 80, 128, 0, R, 1, 0}},
  };
  { size_t n0 = cases.size(); for (size_t i = 0; i < n0; ++i) { Named x = cases[i]; x.c.mode = 1; cases.push_back(x); } }
  for (int grid : {1}) {
    printf("---- grid %d\n", grid);
    for (auto& nc : cases) {
      for (int it = 0; it < 2; ++it) {
        if (nc.c.mode) mb_kernel<1><<<grid, 128, 200 * 1024>>>(nc.c, d); else mb_kernel<0><<<grid, 128, 200 * 1024>>>(nc.c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", nc.name, cudaGetErrorString(e)); return 1; }
      }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const int nm = 3 * (nc.c.K / 16) * nc.c.reps;
      printf("mode %d %-36s mmas/gemm %2d  issue %7.1f clk/gemm (%5.1f/mma)  done %7.1f clk/gemm (%5.1f/mma)\n", nc.c.mode, nc.name,
             3 * nc.c.K / 16, (double)h[0] / nc.c.reps, (double)h[0] / nm, (double)h[1] / nc.c.reps, (double)h[1] / nm);
    }
  }
  return 0;
}
