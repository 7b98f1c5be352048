// fiber_b200 — shared definitions of the small-sequence attention kernels (forward + backward).
//
// Two addressing modes cover every attention on the FIBER path:
//   WINDOW  Swin W-MSA / SW-MSA (swin_transformer.py:195-224 + :363-387): queries/keys are the
//           ws*ws tokens of one window, gathered straight from the image-ordered packed QKV
//           activation through the closed-form cyclic-shift/partition map (no roll / partition /
//           reverse copies exist); relative-position bias and the -100 shift mask are generated
//           in-kernel from the (2ws-1)^2 table and region ids.
//   PLAIN   RoBERTa self-attention (roberta.py:256-326), text->image cross attention (t2i, no
//           mask) and image->text cross attention (i2t, swin_transformer.py:226-259; additive
//           0/-10000 key mask per sample).  Optional attention-probability dropout.
#pragma once
#include "common.cuh"

namespace fiber {

struct AttnParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  float* lse;  // [G, nH, Lq] natural-log sum-exp of the scaled, biased scores
  long long ldq, ldk, ldv, ldo;
  int mode;  // 0 plain, 1 window
  int G;     // plain: batch; window: number of images
  int nH, Lq, Lk;
  float scale;
  const float* key_mask;  // plain: [G, Lk] additive, nullable
  int H, W, ws, shift;    // window geometry (tokens)
  const float* bias_table;  // window: [(2ws-1)^2, nH] fp32
  float drop_p;
  unsigned long long seed;
  // backward only
  const bf16* d_o;
  bf16* dq;
  bf16* dk;
  bf16* dv;
  long long lddo, lddq, lddk, lddv;
  float* dbias_table;  // window: [(2ws-1)^2, nH] fp32, accumulated with atomics
};

#ifdef __CUDACC__
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// exp(x) as one FMUL + one MUFU.EX2 (no denormal range fix-up: results below 2^-126 flush to zero,
// which is what a softmax probability that small rounds to in bf16 anyway).
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
// Counter-based keep decision for attention-probability dropout (same in forward and backward).
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, float p) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<float>(static_cast<uint32_t>(z >> 40)) * (1.0f / 16777216.0f) >= p;
}
#endif

constexpr int ATT_KCHUNK = 48;    // keys per register tile (6 n-tiles of 8)
constexpr int ATT_SKEYS = 144;    // keys staged in shared memory at a time
constexpr int ATT_MAXTOK = 336;   // largest padded window (18x18 = 324 -> 336)
constexpr int ATT_MAXTBL = 1225;  // (2*18-1)^2

}  // namespace fiber
