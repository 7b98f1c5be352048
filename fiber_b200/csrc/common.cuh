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
int option_pdl();  // 1: launch with programmatic stream serialization (FIBER_PDL / fiber_set_option("pdl")); default 0

#ifdef __CUDACC__
// Every kernel of the library is launched through launch_k.  With the "pdl" option on, the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization: the CTAs of this kernel may become resident while the previous
// kernel of the stream is still draining, run their prologue (barrier init, TMEM allocation, tensor-map prefetch,
// shared-memory tables) and then block in pdl_wait() until the previous grid has completed and its writes are visible.
// Contract for every kernel: pdl_wait() is executed unconditionally by every thread BEFORE the first access to global
// memory (read or write) and before any early return; pdl_trigger() sits at the top, so the next kernel's CTAs may
// take the slots this grid frees.  Both instructions are no-ops for a launch without the attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = option_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// Same for a kernel that runs as 2-CTA clusters (CTA pairs on one TPC: tcgen05 cta_group::2); never with PDL.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_pair(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---- programmatic dependent launch (see launch_k) -------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
// After a few failed polls the warp backs off with nanosleep: a spinning warp (the MMA issuer waiting for the epilogue,
// the producer waiting for a free slot, ...) otherwise takes issue slots from the working warps of its scheduler —
// ncu on the GELU GEMM counted ~35 % of all issued instructions in such loops (profiles/r2_gemm_gelu_cache_ncu.txt).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifndef FIBER_NO_BACKOFF
    if (++spins > 4) __nanosleep(spins > 64 ? 256 : 32);
#else
    ++spins;
#endif
    if (spins > (1u << 22)) {
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
// ---- CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster execute one 256-row tcgen05.mma ------------------------
// Each CTA holds its own 128 rows of A and of the accumulator (same TMEM address in both) and HALF of the B tile; the
// leader (cluster rank 0) issues the MMAs.  TMA loads of both CTAs signal the leader's mbarrier; commits are multicast
// to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `mask` once the previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// TMA load into this CTA's smem whose completion bytes go to an mbarrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
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
// Exact (erf-form) GELU as timm's nn.GELU and HF ACT2FN["gelu"] compute it:
//   gelu(x) = x Phi(x),  gelu'(x) = Phi(x) + x phi(x),  Phi(x) = 1/2 + sign(x) (1 - erfc(|x| / sqrt 2)) / 2,
//   phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
// with erfc(u) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-u^2), t = 1 / (1 + p u)   (Abramowitz & Stegun 7.1.26,
// |error| <= 1.5e-7).  The exponential of the erfc and of the density is the SAME number, exp(-x^2 / 2): GELU and
// GELU' together cost one rcp, one ex2 and 16 packed fp32x2 operations per pair of elements (the previous 7.1.28 form
// — a sixth-degree polynomial raised to the 16th power, plus a separate exponential for the density — needed 24), and
// the measured error is smaller (4.6e-7 vs 8.3e-7 absolute on gelu over [-12, 12]).  Every variant below evaluates the
// same operations in the same order, scalar or packed, so all GELU epilogues agree bit for bit.
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
constexpr float GELU_P = 0.3275911f * 0.70710678118654752f;  // p / sqrt 2: t = 1 / (1 + p |x| / sqrt 2)
constexpr float GELU_A1 = 0.254829592f, GELU_A2 = -0.284496736f, GELU_A3 = 1.421413741f, GELU_A4 = -1.453152027f,
                GELU_A5 = 1.061405429f;
constexpr float GELU_C = -0.72134752044448170368f;  // -log2(e) / 2: exp(-x^2 / 2) = 2^(C x^2)
constexpr float GELU_D = 0.39894228040143267794f;   // 1 / sqrt(2 pi)
// Phi(x) and e = exp(-x^2 / 2)
__device__ __forceinline__ void gelu_phi_e(float x, float& phi_cdf, float& e) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(ax, GELU_P, 1.0f));
  float q = fmaf(t, GELU_A5, GELU_A4);
  q = fmaf(q, t, GELU_A3);
  q = fmaf(q, t, GELU_A2);
  q = fmaf(q, t, GELU_A1);
  q *= t;
  e = ex2_approx((x * x) * GELU_C);
  const float r = q * e;  // erfc(|x| / sqrt 2)
  phi_cdf = fmaf(copysignf(0.5f, x), fmaf(r, -1.0f, 1.0f), 0.5f);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float c, e;
  gelu_phi_e(x, c, e);
  return x * c;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float c, e;
  gelu_phi_e(x, c, e);
  return fmaf(x * GELU_D, e, c);
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
// Phi(x) and e = exp(-x^2 / 2) on FOUR packed pairs in lockstep.  The source is written "vertically" — one step across
// all four pairs before the next — so the dependent FFMA2 / FMUL2 / MUFU chains of the pairs interleave instead of
// serialising on their latencies.
__device__ __forceinline__ void gelu_phi_e2x4(const uint64_t (&x)[4], uint64_t (&cdf)[4], uint64_t (&e)[4]) {
  uint64_t t[4], q[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    upk2(x[i], x0, x1);
    t[i] = pk2(fabsf(x0), fabsf(x1));
    sh[i] = pk2(copysignf(0.5f, x0), copysignf(0.5f, x1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(t[i], FIBER_PK2C(GELU_P), FIBER_PK2C(1.0f));
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(x[i], x[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t0, t1;
    upk2(t[i], t0, t1);
    t[i] = pk2(rcp_approx(t0), rcp_approx(t1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(e[i], FIBER_PK2C(GELU_C));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(t[i], FIBER_PK2C(GELU_A5), FIBER_PK2C(GELU_A4));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float e0, e1;
    upk2(e[i], e0, e1);
    e[i] = pk2(ex2_approx(e0), ex2_approx(e1));  // exp(-x^2 / 2)
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], FIBER_PK2C(GELU_A3));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], FIBER_PK2C(GELU_A2));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], FIBER_PK2C(GELU_A1));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = mul2(q[i], t[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = mul2(q[i], e[i]);  // erfc(|x| / sqrt 2)
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], FIBER_PK2C(-1.0f), FIBER_PK2C(1.0f));
#pragma unroll
  for (int i = 0; i < 4; ++i) cdf[i] = fma2(sh[i], q[i], FIBER_PK2C(0.5f));
}
// x[i] <- gelu(x[i]) on four packed pairs
__device__ __forceinline__ void gelu_erf2x4(uint64_t (&x)[4]) {
  uint64_t c[4], e[4];
  gelu_phi_e2x4(x, c, e);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = mul2(x[i], c[i]);
}
// g[i] *= gelu'(x[i]) on four packed pairs
__device__ __forceinline__ void gelu_erf_grad_mul2x4(uint64_t (&g)[4], const uint64_t (&x)[4]) {
  uint64_t c[4], e[4], xd[4];
  gelu_phi_e2x4(x, c, e);
#pragma unroll
  for (int i = 0; i < 4; ++i) xd[i] = mul2(x[i], FIBER_PK2C(GELU_D));
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = fma2(xd[i], e[i], c[i]);  // Phi + x phi
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = mul2(g[i], c[i]);
}
// x[i] <- gelu(x[i]) and g[i] <- gelu'(x[i]) on four packed pairs from ONE evaluation of (Phi, e) (the same operations in
// the same order as gelu_erf2x4 / gelu_erf_grad_mul2x4, so both results are bit-identical to the separate functions)
__device__ __forceinline__ void gelu_erf_both2x4(uint64_t (&x)[4], uint64_t (&g)[4]) {
  uint64_t c[4], e[4], xd[4];
  gelu_phi_e2x4(x, c, e);
#pragma unroll
  for (int i = 0; i < 4; ++i) xd[i] = mul2(x[i], FIBER_PK2C(GELU_D));
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = fma2(xd[i], e[i], c[i]);  // Phi + x phi
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = mul2(x[i], c[i]);
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
