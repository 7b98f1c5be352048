// fiber_b200 — shared device/host helpers for the sm_100a kernels.
// Inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA/TMEM) and
// a few vectorised load/store helpers.  Everything here is B200 (sm_100a) only.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace fiber {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define FIBER_CHECK(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::fiber::set_last_error(__VA_ARGS__);    \
      return -1;                               \
    }                                          \
  } while (0)

#define FIBER_CUDA(expr)                                                                     \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::fiber::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                              __FILE__, __LINE__);                                           \
      return -2;                                                                             \
    }                                                                                        \
  } while (0)

int num_sms();  // SM count of the current device (cached)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("fiber_b200: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 4-D tiled loads / stores (window attention: box = 32 channels x 6 x 6 tokens of a [B, H, W, C] activation)
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/f16 inputs, f32 accumulate, single CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread t of the warp gets lane (base_lane+t), cols c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor (sm_100 format, version 1), SWIZZLE_128B.
//   K-major : rows of 64 bf16 (128 B), 8-row groups 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major: k-rows of 64 contiguous MN elements (128 B); 8 k-rows = one 1024 B swizzle atom
//             (SBO = 1024 between k groups); next 64-wide MN chunk `lbo_bytes` further on.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor: bf16 x bf16 -> f32, dense, M x N, operand majors (0 = K, 1 = MN).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- misc math / memory --------------------------------------------------------------------
// Exact (erf-form) GELU as timm's nn.GELU and HF ACT2FN["gelu"] compute it, with
// erf(u) = 1 - 1 / (1 + a1 u + ... + a6 u^6)^16, u >= 0   (Abramowitz & Stegun 7.1.28, |error| <= 3e-7).
// u = |x| / sqrt(2) is folded into the coefficients; one MUFU reciprocal, no branches.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// returns r = 1 - erf(|x| / sqrt 2) = erfc(|x| / sqrt 2)  in (0, 1]
__device__ __forceinline__ float erfc_abs_scaled(float ax) {
  constexpr float c1 = 0.0705230784f * 0.70710678118654752f;
  constexpr float c2 = 0.0422820123f * 0.5f;
  constexpr float c3 = 0.0092705272f * 0.35355339059327376f;
  constexpr float c4 = 0.0001520143f * 0.25f;
  constexpr float c5 = 0.0002765672f * 0.17677669529663688f;
  constexpr float c6 = 0.0000430638f * 0.125f;
  float t = fmaf(ax, c6, c5);
  t = fmaf(ax, t, c4);
  t = fmaf(ax, t, c3);
  t = fmaf(ax, t, c2);
  t = fmaf(ax, t, c1);
  t = fmaf(ax, t, 1.0f);
  t *= t; t *= t; t *= t; t *= t;
  return rcp_approx(t);
}
// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) = 0.5 (x + |x| (1 - r))
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float r = erfc_abs_scaled(ax);
  const float m = fmaf(-ax, r, ax);
  return fmaf(0.5f, x, 0.5f * m);
}
// gelu'(x) = Phi(x) + x phi(x),  Phi(x) = 0.5 + 0.5 sign(x) (1 - r),  phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float ax = fabsf(x);
  const float r = erfc_abs_scaled(ax);
  const float cdf = fmaf(0.5f, copysignf(1.0f - r, x), 0.5f);
  const float pdf = 0.39894228040143267794f * ex2_approx(x * x * -0.72134752044448170368f);
  return fmaf(x, pdf, cdf);
}
// ---- packed fp32x2 math (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#define FIBER_PK2C(c) ::fiber::pk2((c), (c))
// erfc(|x| / sqrt 2) on FOUR packed pairs in lockstep (same polynomial as erfc_abs_scaled).  The source is
// written "vertically" — one Horner step across all four pairs before the next — so the dependent
// FFMA2 / FMUL2 / MUFU chains of the pairs interleave instead of serialising on their latencies.
__device__ __forceinline__ void erfc_abs_scaled2x4(const uint64_t (&ax)[4], uint64_t (&r)[4]) {
  constexpr float c1 = 0.0705230784f * 0.70710678118654752f;
  constexpr float c2 = 0.0422820123f * 0.5f;
  constexpr float c3 = 0.0092705272f * 0.35355339059327376f;
  constexpr float c4 = 0.0001520143f * 0.25f;
  constexpr float c5 = 0.0002765672f * 0.17677669529663688f;
  constexpr float c6 = 0.0000430638f * 0.125f;
  uint64_t t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], FIBER_PK2C(c6), FIBER_PK2C(c5));
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], t[i], FIBER_PK2C(c4));
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], t[i], FIBER_PK2C(c3));
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], t[i], FIBER_PK2C(c2));
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], t[i], FIBER_PK2C(c1));
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(ax[i], t[i], FIBER_PK2C(1.0f));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = mul2(t[i], t[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t0, t1;
    upk2(t[i], t0, t1);
    r[i] = pk2(rcp_approx(t0), rcp_approx(t1));
  }
}
// gelu on four packed pairs: 0.5 x + |x| (0.5 - 0.5 r)
__device__ __forceinline__ void gelu_erf2x4(uint64_t (&x)[4]) {
  uint64_t ax[4], r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    upk2(x[i], x0, x1);
    ax[i] = pk2(fabsf(x0), fabsf(x1));
  }
  erfc_abs_scaled2x4(ax, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma2(r[i], FIBER_PK2C(-0.5f), FIBER_PK2C(0.5f));
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = mul2(ax[i], r[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = fma2(x[i], FIBER_PK2C(0.5f), r[i]);
}
// g[i] *= gelu'(x[i]) on four packed pairs
__device__ __forceinline__ void gelu_erf_grad_mul2x4(uint64_t (&g)[4], const uint64_t (&x)[4]) {
  uint64_t ax[4], r[4], sh[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    upk2(x[i], x0, x1);
    ax[i] = pk2(fabsf(x0), fabsf(x1));
    sh[i] = pk2(copysignf(0.5f, x0), copysignf(0.5f, x1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(x[i], x[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(e[i], FIBER_PK2C(-0.72134752044448170368f));
  erfc_abs_scaled2x4(ax, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float e0, e1;
    upk2(e[i], e0, e1);
    e[i] = pk2(ex2_approx(e0), ex2_approx(e1));  // exp(-x^2 / 2)
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma2(r[i], FIBER_PK2C(-1.0f), FIBER_PK2C(1.0f));   // 1 - r
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma2(sh[i], r[i], FIBER_PK2C(0.5f));               // Phi(x)
#pragma unroll
  for (int i = 0; i < 4; ++i) ax[i] = mul2(x[i], FIBER_PK2C(0.39894228040143267794f));
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma2(ax[i], e[i], r[i]);                           // Phi + x phi
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = mul2(g[i], r[i]);
}
// x[i] <- gelu(x[i]) and g[i] <- gelu'(x[i]) on four packed pairs from ONE erfc evaluation (the same operation order
// as gelu_erf2x4 / gelu_erf_grad_mul2x4, so both results are bit-identical to the separate functions)
__device__ __forceinline__ void gelu_erf_both2x4(uint64_t (&x)[4], uint64_t (&g)[4]) {
  uint64_t ax[4], r[4], sh[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    upk2(x[i], x0, x1);
    ax[i] = pk2(fabsf(x0), fabsf(x1));
    sh[i] = pk2(copysignf(0.5f, x0), copysignf(0.5f, x1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(x[i], x[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(e[i], FIBER_PK2C(-0.72134752044448170368f));
  erfc_abs_scaled2x4(ax, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float e0, e1;
    upk2(e[i], e0, e1);
    e[i] = pk2(ex2_approx(e0), ex2_approx(e1));  // exp(-x^2 / 2)
  }
  uint64_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = fma2(r[i], FIBER_PK2C(-0.5f), FIBER_PK2C(0.5f));   // 0.5 (1 - r)
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma2(r[i], FIBER_PK2C(-1.0f), FIBER_PK2C(1.0f));   // 1 - r
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = fma2(sh[i], r[i], FIBER_PK2C(0.5f));               // Phi(x)
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = mul2(x[i], FIBER_PK2C(0.39894228040143267794f));
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = fma2(r[i], e[i], g[i]);                            // Phi + x phi
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = mul2(ax[i], h[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = fma2(x[i], FIBER_PK2C(0.5f), h[i]);                // 0.5 x + |x| 0.5 (1 - r)
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(t);
}

#endif  // __CUDACC__

}  // namespace fiber
