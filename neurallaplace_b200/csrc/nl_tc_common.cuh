// tcgen05 / TMEM / mbarrier primitives (inline PTX) and the shared-memory operand layout used by
// the tensor-core fused kernels.  sm_100a only.
//
// Operand layout ("chunk-major", no swizzle): a matrix X[rows][cols] of bf16 is stored as
//     byte_offset(r, c) = (c / 8) * (ROWS * 16) + r * 16 + (c % 8) * 2
// i.e. [cols/8][ROWS][8 bf16].  One thread owning row r writes 16-byte pieces -> consecutive
// threads hit consecutive 16-byte words (conflict-free, vectorised).  The SAME bytes are a valid
// canonical UMMA operand in two ways (descriptor layout type SWIZZLE_NONE):
//   * K-major  (MN = r, K = c):  core matrix = 8 rows x 16 B;  LBO = ROWS*16 (next K chunk),
//                                SBO = 128 (next 8-row group);
//   * MN-major (MN = c, K = r):  core matrix = 8 k x 16 B;     SBO = ROWS*16 (next 8 MN values),
//                                LBO = 128 (next 8-k group).
// so an activation tile feeds both the forward-type GEMM (reduction over its columns) and the
// weight-gradient GEMM (reduction over its rows) without a transpose.
//
// fp32 parity: every operand is split v = hi + lo with hi = bf16(v), lo = bf16(v - hi) and each GEMM
// runs 3 passes hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM ("bf16x3"): relative error
// ~2^-16 per product, measured 1e-5 on x(t) and all gradients (tolerance 1e-4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nl_pack2.cuh"

namespace nl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp (deterministic for a given member mask).  tcgen05.mma must be issued from a
// branch on THIS predicate: ptxas then knows the region is single-lane and keeps the descriptors in uniform
// registers.  Behind a plain `tid == 0` test it wraps every UTCHMMA in a per-lane "waterfall" loop (ELECT +
// BRA.U.ANY + 4 R2UR), measured at ~90 clk per MMA instead of ~50 (scripts/mb/tc_mb.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait: a protocol bug must abort the kernel (trap), never hang the GPU.
// try_wait carries a suspend-time hint: without one the hardware returns "not yet" after a short system-dependent time
// and the loop polls -- ncu counted 39 M try_wait warp-instructions per launch of the two-tile saved backward (1 200 per
// tile, ~20 % of its issued instructions with the loop around them), each of them a shared-memory wavefront competing
// with the tensor core's operand reads (the kernel is shared-memory-bandwidth bound).  With the hint the warp sleeps in
// hardware until the phase completes.
#ifndef NL_MBAR_HINT_NS
#define NL_MBAR_HINT_NS 20000u
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
#if NL_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(NL_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
    if (done) return;
  }
  __trap();
}

// Same, for the dedicated issuer warps: they wait most of the time, and a hot polling loop would take issue
// slots from the epilogue warps of the same scheduler; back off between polls.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
    // suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint expires
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(100000u)
        : "memory");
    if (done) return;
  }
  __trap();
}

// ---- proxy / tcgen05 fences ----------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp) ---------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  return d;
}
// instruction descriptor: kind::f16, A/B = bf16, D = f32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM loads: lane = 32*(warp%4) + thread-in-warp, N consecutive 32-bit columns ---------------
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- TMEM stores (A operand of a TS-mode MMA): lane = 32*(warp%4) + thread-in-warp, N consecutive columns ----
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]  (A: lane = row, each 32-bit column = two consecutive K elements)
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

// fire-and-forget vector reduction: 4 consecutive floats (16-byte aligned) with ONE LSU operation per lane
// (scalar REDs cost ~1.3 clk per lane each: draining a 128 x 80 accumulator with them took 16 k clk)
__device__ __forceinline__ void red_add4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- bf16 hi/lo split ------------------------------------------------------------------------
// packs (v0, v1) -> hi = bf16x2(v0 | v1<<16), lo = bf16x2 of the residuals
__device__ __forceinline__ void split2(float v0, float v1, uint32_t* hi, uint32_t* lo) {
  uint32_t h;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
  float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xFFFF0000u);
  float r0 = v0 - h0, r1 = v1 - h1;
  uint32_t l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
  *hi = h;
  *lo = l;
}

// store 8 consecutive columns [c0, c0+8) of row r (c0 % 8 == 0) into hi / lo operand images
__device__ __forceinline__ void store_chunk8(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                             int c0, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], &h[i], &l[i]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16;
  *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Three-term split v = hi + mid + lo (bf16 each, 24 mantissa bits: float32 exactly up to its last bit or two).
// (hi, mid) are bit-identical to the two-term (hi, lo) pair.  Used where 2^-17 of operand representation error
// is too much: the De Hoog forward, whose quotient-difference recurrence amplifies noise on F by 10^2..10^3.
__device__ __forceinline__ void split3(float v0, float v1, uint32_t* hi, uint32_t* mid, uint32_t* lo) {
  uint32_t h, m, l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
  const float r0 = v0 - __uint_as_float(h << 16), r1 = v1 - __uint_as_float(h & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(r1), "f"(r0));
  const float q0 = r0 - __uint_as_float(m << 16), q1 = r1 - __uint_as_float(m & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(q1), "f"(q0));
  *hi = h;
  *mid = m;
  *lo = l;
}

// store 8 consecutive columns of row r into the three images img0 + {0, 1, 2} * img_stride
__device__ __forceinline__ void store_chunk8x3(unsigned char* img0, size_t img_stride, int rows_in_image, int r,
                                               int c0, const float* v) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split3(v[2 * i], v[2 * i + 1], &h[i], &m[i], &l[i]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16;
  *reinterpret_cast<uint4*>(img0 + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img0 + img_stride + off) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4*>(img0 + 2 * img_stride + off) = make_uint4(l[0], l[1], l[2], l[3]);
}


// ---- packed variants (values arrive as FP32 pairs, nl_pack2.cuh): the residual subtraction is one FADD2 ----
__device__ __forceinline__ void split2p(p2::f2 v, uint32_t* hi, uint32_t* lo) {
  float v0, v1;
  p2::unpk(v, v0, v1);
  uint32_t h;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
  const p2::f2 r = p2::sub2(v, p2::pk(__uint_as_float(h << 16), __uint_as_float(h & 0xFFFF0000u)));
  float r0, r1;
  p2::unpk(r, r0, r1);
  uint32_t l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
  *hi = h;
  *lo = l;
}
__device__ __forceinline__ void split3p(p2::f2 v, uint32_t* hi, uint32_t* mid, uint32_t* lo) {
  float v0, v1;
  p2::unpk(v, v0, v1);
  uint32_t h, m, l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
  const p2::f2 r = p2::sub2(v, p2::pk(__uint_as_float(h << 16), __uint_as_float(h & 0xFFFF0000u)));
  float r0, r1;
  p2::unpk(r, r0, r1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(r1), "f"(r0));
  const p2::f2 q = p2::sub2(r, p2::pk(__uint_as_float(m << 16), __uint_as_float(m & 0xFFFF0000u)));
  float q0, q1;
  p2::unpk(q, q0, q1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(q1), "f"(q0));
  *hi = h;
  *mid = m;
  *lo = l;
}
// 8 consecutive columns given as four pairs
__device__ __forceinline__ void store_chunk8p(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                              int c0, const p2::f2* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2p(v[i], &h[i], &l[i]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16;
  *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void store_chunk8x3p(unsigned char* img0, size_t img_stride, int rows_in_image, int r,
                                                int c0, const p2::f2* v) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split3p(v[i], &h[i], &m[i], &l[i]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16;
  *reinterpret_cast<uint4*>(img0 + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img0 + img_stride + off) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4*>(img0 + 2 * img_stride + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
// 4 consecutive columns given as two pairs
__device__ __forceinline__ void store_chunk4p(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                              int c0, p2::f2 v0, p2::f2 v1) {
  uint32_t h[2], l[2];
  split2p(v0, &h[0], &l[0]);
  split2p(v1, &h[1], &l[1]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16 + (size_t)(c0 & 7) * 2;
  *reinterpret_cast<uint2*>(img_hi + off) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(img_lo + off) = make_uint2(l[0], l[1]);
}

// store 4 consecutive columns [c0, c0+4) of row r (c0 % 4 == 0): half of a 16-byte piece
__device__ __forceinline__ void store_chunk4(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                             int c0, const float* v) {
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) split2(v[2 * i], v[2 * i + 1], &h[i], &l[i]);
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16 + (size_t)(c0 & 7) * 2;
  *reinterpret_cast<uint2*>(img_hi + off) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(img_lo + off) = make_uint2(l[0], l[1]);
}

// the same in two steps, for callers that also want the packed words (TMEM copy of the operand)
__device__ __forceinline__ void store_packed4(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                              int c0, const uint32_t* h, const uint32_t* l) {
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16 + (size_t)(c0 & 7) * 2;
  *reinterpret_cast<uint2*>(img_hi + off) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(img_lo + off) = make_uint2(l[0], l[1]);
}
__device__ __forceinline__ void store_packed8(unsigned char* img_hi, unsigned char* img_lo, int rows_in_image, int r,
                                              int c0, const uint32_t* h, const uint32_t* l) {
  const size_t off = (size_t)(c0 >> 3) * ((size_t)rows_in_image * 16) + (size_t)r * 16;
  *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// bytes of one operand image (hi or lo) holding [rows][cols] bf16, cols % 8 == 0
__host__ __device__ constexpr size_t image_bytes(int rows, int cols) { return (size_t)rows * cols * 2; }

// One GEMM = 3 passes (hi*hi, hi*lo, lo*hi) x ksteps MMAs.  Addresses are 32-bit shared addresses of
// the hi images; *_lo_off is the byte distance from the hi image to the lo image.
// a_step / b_step: byte advance of the descriptor start per K step of 16 elements.
struct GemmOp {
  uint32_t a_hi, a_lo_off, a_lbo, a_sbo, a_step;
  uint32_t b_hi, b_lo_off, b_lbo, b_sbo, b_step;
  uint32_t idesc;
  int ksteps;
};

#ifndef NL_TC_PASS_MODE
#define NL_TC_PASS_MODE 0
#endif
__device__ __forceinline__ void issue_gemm(const GemmOp& g, uint32_t d_tmem, bool accumulate) {
  // descriptors are built once per pass; a K step only bumps the 14-bit start-address field
  const uint64_t a_inc = (uint64_t)(g.a_step >> 4), b_inc = (uint64_t)(g.b_step >> 4);
  uint32_t acc = accumulate ? 1u : 0u;
  // NL_TC_PASS_MODE (experiments): 0 = hi.hi, hi.lo, lo.hi ; 1 = lo.hi, hi.lo, hi.hi ; 2 = lo.lo, lo.hi, hi.lo, hi.hi
  constexpr int kPasses = NL_TC_PASS_MODE == 2 ? 4 : 3;
#pragma unroll 1
  for (int pass = 0; pass < kPasses; ++pass) {
    int a_lo, b_lo;
    if (NL_TC_PASS_MODE == 0) { a_lo = pass == 2; b_lo = pass == 1; }
    else if (NL_TC_PASS_MODE == 1) { a_lo = pass == 0; b_lo = pass == 1; }
    else { a_lo = pass <= 1; b_lo = pass == 0 || pass == 2; }
    uint64_t ad = make_sdesc(g.a_hi + (a_lo ? g.a_lo_off : 0), g.a_lbo, g.a_sbo);
    uint64_t bd = make_sdesc(g.b_hi + (b_lo ? g.b_lo_off : 0), g.b_lbo, g.b_sbo);
#pragma unroll 4
    for (int ks = 0; ks < g.ksteps; ++ks) {
      umma_bf16(d_tmem, ad, bd, g.idesc, acc);
      acc = 1u;
      ad += a_inc;
      bd += b_inc;
    }
  }
}

// Six passes over three-term operands (images at hi + {0,1,2} * lo_off): every product down to 2^-24 relative
//   hi.hi + (hi.mid + mid.hi) + (hi.lo + lo.hi + mid.mid)      ("bf16x6"; smallest terms first)
__device__ __forceinline__ void issue_gemm6(const GemmOp& g, uint32_t d_tmem, bool accumulate) {
  const uint64_t a_inc = (uint64_t)(g.a_step >> 4), b_inc = (uint64_t)(g.b_step >> 4);
  uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll 1
  for (int pass = 0; pass < 6; ++pass) {
    // (a image, b image): (2,0) (0,2) (1,1) (1,0) (0,1) (0,0)
    const int ai = pass == 0 ? 2 : (pass == 2 || pass == 3) ? 1 : 0;
    const int bi = pass == 1 ? 2 : (pass == 2 || pass == 4) ? 1 : 0;
    uint64_t ad = make_sdesc(g.a_hi + ai * g.a_lo_off, g.a_lbo, g.a_sbo);
    uint64_t bd = make_sdesc(g.b_hi + bi * g.b_lo_off, g.b_lbo, g.b_sbo);
#pragma unroll 4
    for (int ks = 0; ks < g.ksteps; ++ks) {
      umma_bf16(d_tmem, ad, bd, g.idesc, acc);
      acc = 1u;
      ad += a_inc;
      bd += b_inc;
    }
  }
}

// Same three passes with the A operand in TMEM (hi image at a_hi, lo image at a_lo, 8 columns per K step of 16):
// the dispatch no longer streams 4 KB of A through the shared-memory port (measured 32 instead of 50 clk).
__device__ __forceinline__ void issue_gemm_ts(const GemmOp& g, uint32_t a_hi, uint32_t a_lo, uint32_t d_tmem,
                                              bool accumulate) {
  const uint64_t b_inc = (uint64_t)(g.b_step >> 4);
  uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    uint32_t at = pass == 2 ? a_lo : a_hi;
    uint64_t bd = make_sdesc(g.b_hi + (pass == 1 ? g.b_lo_off : 0), g.b_lbo, g.b_sbo);
#pragma unroll 4
    for (int ks = 0; ks < g.ksteps; ++ks) {
      umma_bf16_ts(d_tmem, at, bd, g.idesc, acc);
      acc = 1u;
      at += 8;
      bd += b_inc;
    }
  }
}

// K-major use of an image [rows][cols]: reduction over cols, 16 per step
__device__ __forceinline__ void operand_kmajor(uint32_t img, uint32_t lo_off, int rows, uint32_t* base,
                                               uint32_t* lo, uint32_t* lbo, uint32_t* sbo, uint32_t* step) {
  *base = img; *lo = lo_off; *lbo = rows * 16; *sbo = 128; *step = 2 * rows * 16;
}
// MN-major use of the same image: reduction over rows, 16 per step
__device__ __forceinline__ void operand_mnmajor(uint32_t img, uint32_t lo_off, int rows, uint32_t* base,
                                                uint32_t* lo, uint32_t* lbo, uint32_t* sbo, uint32_t* step) {
  *base = img; *lo = lo_off; *lbo = 128; *sbo = rows * 16; *step = 256;
}

}  // namespace tc
}  // namespace nl
